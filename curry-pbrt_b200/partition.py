"""Sample-index partition of render() across GPUs.

Rank g of G renders a contiguous range of the sample indices of EVERY pixel -- the same rule
cpbrt_render applies when a communicator is attached and no explicit range is given
(cpbrt_sample_share in csrc/host/host_api.cpp; tests compare the two).  A rank that starts at sample s
first has to replay the film-point drift of the s samples before it (two radical inverses each,
render.rs:26-31), which costs about REPLAY_COST of rendering a sample.  Equal shares would make the
last rank the slowest, so the shares shrink geometrically: n_g ~ (1 - REPLAY_COST)^g, i.e.
rendering n_g samples plus replaying the s_g before them takes the same time on every rank.
Pure integer / IEEE-double logic, importable without the CUDA library.
"""

REPLAY_COST = 1.0 / 420.0     # measured on the Cornell box (round 2): 2.06 ms to replay 899 samples (k_film_replay), 0.97 ms to render one


def _boundary(spp, k, n_ranks):
    """First sample index of rank k (k = n_ranks: spp).  Same operation order as the C side."""
    if k <= 0:
        return 0
    if k >= n_ranks:
        return spp
    q = 1.0 - REPLAY_COST
    qk, qn = 1.0, 1.0
    for i in range(n_ranks):
        qn *= q
        if i < k:
            qk *= q
    b = int(spp * ((1.0 - qk) / (1.0 - qn)) + 0.5)
    return min(max(b, 0), spp)


def sample_share(spp, rank, n_ranks):
    return _boundary(spp, rank, n_ranks), _boundary(spp, rank + 1, n_ranks)
