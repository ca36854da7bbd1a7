"""Host-side score normalisation (drop-in for /root/reference/utils/metric_transformations.py:3-38).
Four scalars in, four scalars out; f64 numpy, no kernel."""
import numpy as np


def transform_metrics(path_length_similarity, trajectory_mse, directional_consistency, distribution_similarity):
    """Map the four headline metrics to [0, 1] scores.

    path length similarity passes through; trajectory MSE is clipped at 0, log1p'd, scaled by
    log1p(1) and inverted; directional consistency is folded with abs(); distribution similarity
    is log1p'd and scaled by log1p(1).  Keys match the heat-map consumer
    (scripts/analysis/analyze_trajectory_metrics.py:90-100)."""
    unit = np.log1p(1.0)
    mse_score = np.clip(1 - np.log1p(np.clip(trajectory_mse, 0, None)) / unit, 0, 1)
    dist_score = np.clip(np.log1p(distribution_similarity) / unit, 0, 1)
    return {
        "path_length_similarity": path_length_similarity,
        "trajectory_mse": mse_score,
        "mean_directional_consistency": np.abs(directional_consistency),
        "distribution_similarity": dist_score,
    }
