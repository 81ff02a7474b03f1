"""Sample-index partition of render() across GPUs.

Rank g of G renders samples [spp*g/G, spp*(g+1)/G) of EVERY pixel -- the same rule
cpbrt_render applies when a communicator is attached and no explicit range is given
(csrc/cpbrt_api.cu).  Pure integer logic, importable without the CUDA library.
"""


def sample_share(spp, rank, n_ranks):
    return spp * rank // n_ranks, spp * (rank + 1) // n_ranks
