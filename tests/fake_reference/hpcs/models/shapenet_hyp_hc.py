from hpcs import ReferencePathReached
from hpcs.models.base_hyp_hc import BaseSimilarityHypHC
from hpcs.utils.data import to_categorical


class ShapeNetHypHC(BaseSimilarityHypHC):
    def __init__(self, *args, train_rotation='so3', test_rotation='so3', class_vector=False, **kwargs):
        super().__init__(*args, **kwargs)
        self.num_categories = 16
        self.train_rotation, self.test_rotation, self.class_vector = train_rotation, test_rotation, class_vector

    def _forward(self, batch, testing):
        raise ReferencePathReached("ShapeNetHypHC._forward (host rotation)")
