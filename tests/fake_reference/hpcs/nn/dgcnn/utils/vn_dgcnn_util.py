from hpcs import unpatched

knn = unpatched("hpcs.nn.dgcnn.utils.vn_dgcnn_util.knn")
get_graph_feature = unpatched("hpcs.nn.dgcnn.utils.vn_dgcnn_util.get_graph_feature")
get_graph_feature_cross = unpatched("hpcs.nn.dgcnn.utils.vn_dgcnn_util.get_graph_feature_cross")
