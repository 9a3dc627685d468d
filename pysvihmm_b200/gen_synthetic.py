"""Synthetic HMM data, mirroring gen_synthetic.py of the reference (input generator for parity:
draws come from the legacy global numpy RNG in the same order, so a seed reproduces the
reference's series bit for bit)."""
import numpy as np

from .util import make_mask, make_mask_prediction  # noqa: F401


def generate_data(tran, emit, T, miss=0., nmasks=1):
    """gen_synthetic.py:8-56.  emit[k] must offer rvs() like pybasicbayes Gaussian
    (distributions.py:123-126).  Returns obs (T,D), sts (T,), masks."""
    K = tran.shape[0]
    curr_st = 0
    sts = [0]
    obs = [emit[0].rvs()[0]]
    for _ in range(T - 1):
        curr_st = np.random.choice(K, p=tran[curr_st, :])
        sts.append(curr_st)
        obs.append(emit[curr_st].rvs()[0])
    obs = np.array(obs)
    sts = np.array(sts)
    masks = None
    if miss > 0.:
        masks = [make_mask(sts, miss) for _ in range(nmasks)]
        if len(masks) == 1:
            masks = masks[0]
    return obs, sts, masks


def _chain(tran, emit, T):
    """The sampling loop shared by gen_synthetic.py:8-56, :59-107 and :110-155 (same draw order)."""
    K = tran.shape[0]
    curr_st = 0
    sts = [0]
    obs = [emit[0].rvs()[0]]
    for _ in range(T - 1):
        curr_st = np.random.choice(K, p=tran[curr_st, :])
        sts.append(curr_st)
        obs.append(emit[curr_st].rvs()[0])
    return np.array(obs), np.array(sts)


def generate_data_smoothing(tran, emit, T, miss=0., left=0, nmasks=1):
    """gen_synthetic.py:59-107: a `miss` fraction of the data right of index `left` is missing."""
    obs, sts = _chain(tran, emit, T)
    masks = None
    if miss > 0.:
        masks = [make_mask(sts, miss, left) for _ in range(nmasks)]
        if len(masks) == 1:
            masks = masks[0]
    return obs, sts, masks


def generate_data_prediction(tran, emit, T, miss=0., nmasks=1):
    """gen_synthetic.py:110-155: the last `miss` fraction of the sequence is missing."""
    obs, sts = _chain(tran, emit, T)
    masks = None
    if miss > 0:
        masks = [make_mask_prediction(sts, miss) for _ in range(nmasks)]
        if len(masks) == 1:
            masks = masks[0]
    return obs, sts, masks


def read_data_chunks(fname, T, D, size):
    """gen_synthetic.py:188-191 (read_data_mmap): generator over chunks of `size` rows of a float64
    memmap -- the feeder for series larger than host memory; a chunk (or the whole memmap via
    read_data_mmap) goes to EStepEngine.set_series / set_series_streamed."""
    fp = np.memmap(fname, dtype='float64', mode='r', shape=(T, D))
    for i in range(T // size):
        yield np.array(fp[i * size:(i + 1) * size, :])


def generate_data_mmap(tran, emit, T, fname, chunk=100000):
    """gen_synthetic.py:158-185: stream a long series into a float64 memmap on disk."""
    D = len(emit[0].mu)
    mm = np.memmap(fname, dtype='float64', mode='w+', shape=(T, D))
    K = tran.shape[0]
    curr_st = 0
    sts = np.empty(T, dtype=np.int64)
    for t in range(T):
        if t > 0:
            curr_st = np.random.choice(K, p=tran[curr_st, :])
        sts[t] = curr_st
        mm[t] = emit[curr_st].rvs()[0]
        if t % chunk == 0:
            mm.flush()
    mm.flush()
    return sts


def read_data_mmap(fname, T, D):
    """gen_synthetic.py:188-191."""
    return np.memmap(fname, dtype='float64', mode='r', shape=(T, D))
