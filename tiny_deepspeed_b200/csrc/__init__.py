"""Native sm_100a sources (CUDA + C++) and their in-tree build driver."""
