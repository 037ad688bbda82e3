"""Model of the slot bookkeeping of the compacted row copies in conv_fwd_tc_kernel<.., kCompact = true> (csrc/conv_tc.cu):
a ring slot keeps a 128-bit mask of rows that hold something other than zeros; a fill touches only rows that are valid now
(data copy, by the lanes of the slice's valid 16-byte chunks) or dirty from the slot's previous use (zero fill, by all
eight chunk lanes).  The invariant the tensor core relies on: after every fill, the chunks the MMA reads hold the
gathered data for valid rows and zeros for all others -- also when full-width and partial last slices alternate in a
slot.  (The kernel itself needs a GPU; this pins the rule it implements.)"""
import numpy as np
import pytest


def simulate(n_slots, nq, last_chunks, n_iters, p_valid, seed):
    rng = np.random.default_rng(seed)
    slot_data = rng.integers(1, 1 << 30, (n_slots, 128, 8))          # uninitialised shared memory: garbage
    dirty = np.ones((n_slots, 128), bool)                            # unknown content: every row counts as dirty
    touched = 0
    for g in range(n_iters):
        s, q = g % n_slots, g % nq                                   # PA == stages: iteration g always lands in slot g mod S
        width = 8 if q + 1 < nq else last_chunks                     # 16-byte chunks the slice really has
        valid = rng.random(128) < p_valid
        tag = rng.integers(1, 1 << 30, (128, 8))                     # what the gathered rows contain
        touch = valid | dirty[s]
        touched += int(touch.sum())
        for r in np.nonzero(touch)[0]:
            if valid[r]:
                slot_data[s, r, :width] = tag[r, :width]             # data copy: only lanes of valid chunks issue
            else:
                slot_data[s, r, :] = 0                               # zero fill: all eight chunk lanes issue
        dirty[s] = valid
        # the MMA of this stage reads chunks [0, width)
        expect = np.where(valid[:, None], tag[:, :width], 0)
        np.testing.assert_array_equal(slot_data[s, :, :width], expect)
    return touched / (128 * n_iters)


@pytest.mark.parametrize("n_slots,nq,last_chunks", [(7, 2, 4), (8, 1, 8), (8, 1, 2), (6, 3, 8), (4, 6, 4), (7, 2, 8), (5, 4, 6)])
@pytest.mark.parametrize("p_valid", [0.0, 0.2, 0.34, 1.0])
def test_slot_invariant(n_slots, nq, last_chunks, p_valid):
    frac = simulate(n_slots, nq, last_chunks, 400, p_valid, seed=n_slots * 100 + nq)
    if 0.0 < p_valid < 1.0:
        assert frac < 2 * p_valid + 0.05        # rows touched per stage: at most valid-now + valid-before (plus the first fills)
