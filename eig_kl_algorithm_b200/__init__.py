"""eig_kl_algorithm_b200 -- B200-native (sm_100a) EIG+KL hypergraph bipartitioner.

Drop-in for the cEIG / cKL / gKL path of yhinai/EIG-KL-Algorithm; see DESIGN.md.
"""
__version__ = "0.1.0"
