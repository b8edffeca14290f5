from hpcs.nn.pointnet.utils.vn_dgcnn_util import get_graph_feature_cross
