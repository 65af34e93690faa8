"""umeregrobust_b200 — B200-native (sm_100a) implementation of UMERegRobust's UME
descriptor-and-registration hot path, behind the reference's own Python call signatures.

    from umeregrobust_b200 import patch_reference
    patch_reference()          # evaluate.py / utils.loc_utils now call the CUDA path

See DESIGN.md for the path and its boundary, include/umereg_b200.h for the C ABI.
"""
from .api import (ball_query, knn_points, knn_gather, knn1_transfer, ume_moments, ume_moments_pair, ume_moments_backward,
                  rigid_solve_backward, ume_cdist_backward,
                  neighbor_count, my_ume_generation,
                  create_local_ume_matrix, ume_descriptors, ume_descriptors_split, descriptor_cdist, descriptor_cdist_split, ume_cdist, rigid_solve,
                  batch_estimate_transform_ume_old, relative_rotation_error, ball_query_gather, ume_kp_layer,
                  register_hypotheses, feature_spatial_var, cauchy_kernel, correlation_scores,
                  pc_corr_cost_pytorch3d, weighted_features, FeatureCorrelator, weighted_match_subsample, sparse_quantize,
                  select_hypothesis, linear_sum_assignment, hungarian_match, config)
from .patch import patch_reference

__all__ = ["ball_query", "knn_points", "knn_gather", "knn1_transfer", "ume_moments", "ume_moments_pair", "ume_moments_backward",
           "rigid_solve_backward", "ume_cdist_backward",
           "neighbor_count", "my_ume_generation",
           "create_local_ume_matrix", "ume_descriptors", "ume_descriptors_split", "descriptor_cdist", "descriptor_cdist_split", "ume_cdist", "rigid_solve",
           "batch_estimate_transform_ume_old", "relative_rotation_error", "ball_query_gather", "ume_kp_layer",
           "register_hypotheses", "feature_spatial_var", "cauchy_kernel", "correlation_scores",
           "pc_corr_cost_pytorch3d", "weighted_features", "FeatureCorrelator", "weighted_match_subsample",
           "sparse_quantize", "select_hypothesis", "linear_sum_assignment", "hungarian_match",
           "patch_reference", "config"]
