"""convex_dim_red -- B200 (sm_100a) build of the alternating-update hot path of
azedarach/matrix-factorization-case-studies.

Same package exports as the reference (``src/convex_dim_red/__init__.py:5-11``);
the numerical work runs in ``libcdr_b200.so`` (see ``include/cdr_b200.h``).
"""

from .archetypal_analysis import ArchetypalAnalysis, KernelAA
from .furthest_sum import furthest_sum
from .gpnh_convex_coding import GPNHConvexCoding
from .kmeans import KMeans, gap_statistic, kmeans_lloyd, kmeans_plusplus
from .pca import PCA
from .simplex_projection import (simplex_project_rows, simplex_project_columns)
from .spg import spg
from .stochastic_matrices import left_stochastic_matrix, right_stochastic_matrix

__all__ = ['ArchetypalAnalysis', 'KernelAA', 'GPNHConvexCoding', 'KMeans', 'PCA', 'furthest_sum',
           'kmeans_plusplus',
           'gap_statistic', 'kmeans_lloyd', 'simplex_project_rows', 'simplex_project_columns',
           'spg', 'left_stochastic_matrix', 'right_stochastic_matrix']
