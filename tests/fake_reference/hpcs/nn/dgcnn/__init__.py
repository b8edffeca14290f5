from .dgcnn_partseg import DGCNN_partseg
from .vn_dgcnn_partseg import VN_DGCNN_partseg
from .vn_dgcnn_expo import VN_DGCNN_expo
