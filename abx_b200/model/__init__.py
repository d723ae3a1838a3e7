"""Score network of the AbX sampler with the reference's module names and state_dict layout
(abx/model/*.py), running its hot ops on the sm_100a kernels of libabx_b200."""
