"""B200-native cluster-registration engine: drop-in for the cluster-ICP hot path of
jl6017/AutoURDF (PointCloud/cluster_icp.py, mlp_reg.py:155-170, dq_func.py).

Operators live in ``libaurdf.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/aurdf.h``); the Python modules mirror the reference's module and function names.
"""
__version__ = "0.1.0"
