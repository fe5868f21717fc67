"""Property tests of the host-side helpers (CPU only): shard arithmetic for any (total, world, align) and the landmark
regressor normalisation for every input form the reference produces (scipy sparse, dense, stacked torch COO)."""
import numpy as np
import scipy.sparse as sp
import torch
from hypothesis import given, settings, strategies as st

from chore_b200 import dist as cdist
from chore_b200.smpl import _to_csr


@settings(max_examples=200, deadline=None)
@given(total=st.integers(0, 5_000_000), world=st.integers(1, 16), align=st.sampled_from([1, 128, 256]))
def test_shard_range_partitions_exactly(total, world, align):
    pos = 0
    for r in range(world):
        start, count = cdist.shard_range(total, r, world, align=align)
        assert start == pos and count >= 0
        assert start % align == 0 or count == 0 or start == total      # shard boundaries fall on tile boundaries
        pos += count
    assert pos == total


@settings(max_examples=50, deadline=None)
@given(batch=st.integers(0, 200), world=st.integers(1, 8))
def test_shard_images_is_a_contiguous_partition(batch, world):
    seen = [i for r in range(world) for i in cdist.shard_images(batch, r, world)]
    assert seen == list(range(batch))
    sizes = [len(cdist.shard_images(batch, r, world)) for r in range(world)]
    assert max(sizes) - min(sizes) <= 1


def test_regressor_forms_agree():
    rng = np.random.default_rng(0)
    dense = (rng.random((25, 300)) < 0.05) * rng.random((25, 300))
    dense[3] = 0.0                                              # an empty row must survive
    a = _to_csr(sp.coo_matrix(dense))
    b = _to_csr(torch.from_numpy(dense))
    coo = sp.coo_matrix(dense)
    t = torch.sparse_coo_tensor(np.vstack([coo.row, coo.col]), coo.data, coo.shape)
    c = _to_csr(torch.stack([t, t, t]))                         # the reference stacks one copy per batch element
    # duplicates are summed (torch_functions.batch_sparse_dense_matmul coalesces)
    dup = sp.coo_matrix((np.r_[coo.data, coo.data[:5]], (np.r_[coo.row, coo.row[:5]], np.r_[coo.col, coo.col[:5]])), shape=coo.shape)
    d = _to_csr(dup)
    want_dup = dense.copy()
    for r, cidx, v in zip(coo.row[:5], coo.col[:5], coo.data[:5]):
        want_dup[r, cidx] += v
    for m, want in ((a, dense), (b, dense), (c, dense), (d, want_dup)):
        assert m.shape == (25, 300) and m.has_sorted_indices
        np.testing.assert_allclose(m.toarray(), want, rtol=0, atol=1e-12)
