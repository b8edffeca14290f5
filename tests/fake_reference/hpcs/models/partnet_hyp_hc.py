from hpcs import ReferencePathReached
from hpcs.models.base_hyp_hc import BaseSimilarityHypHC
from hpcs.utils.data import to_categorical
from hpcs.loss.ultrametric_loss import HierarchicalMetricHyperbolicLoss


class PartNetHypHC(BaseSimilarityHypHC):
    def __init__(self, *args, hierarchical=False, hierarchy_list=(), train_rotation='so3', test_rotation='so3',
                 class_vector=False, **kwargs):
        super().__init__(*args, **kwargs)
        self.train_rotation, self.test_rotation, self.class_vector = train_rotation, test_rotation, class_vector
        self.hierarchical, self.hierarchy_list = hierarchical, hierarchy_list
        if self.hierarchical:
            self.metric_hyp_loss = HierarchicalMetricHyperbolicLoss(
                margin=self.margin, t_per_anchor=self.t_per_anchor, fraction=self.fraction, scale=self.scale,
                temperature=self.temperature, anneal_factor=self.anneal_factor, num_class=self.num_class,
                embedding_size=self.euclidean_size, miner=self.miner, hierarchy_list=self.hierarchy_list)

    def _forward(self, batch, testing):
        raise ReferencePathReached("PartNetHypHC._forward (host rotation)")
