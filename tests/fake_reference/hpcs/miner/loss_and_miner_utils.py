from hpcs import unpatched

get_balanced_random_triplet_indices = unpatched("hpcs.miner.loss_and_miner_utils.get_balanced_random_triplet_indices")
