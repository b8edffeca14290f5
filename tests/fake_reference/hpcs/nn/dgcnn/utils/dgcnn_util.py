from hpcs import unpatched

knn = unpatched("hpcs.nn.dgcnn.utils.dgcnn_util.knn")
get_graph_feature = unpatched("hpcs.nn.dgcnn.utils.dgcnn_util.get_graph_feature")   # [B,C,N] layout: not a target
