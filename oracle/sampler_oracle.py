"""CPU restatement of the on-device patch sampler (madeleine_b200/csrc/sampler.cu) — TEST INFRASTRUCTURE ONLY.

The reference rule it implements is SlideDataset.sample_n (madeleine/datasets/wsi_dataset.py:42-50): a bag with at least
`sample` rows contributes a uniformly random subset WITHOUT replacement (``randperm(N)[:sample]``), a shorter one `sample`
draws WITH replacement (``randint(0, N, (sample,))``), a missing stain a zero bag (wsi_dataset.py:66).  torch's CPU generator
stream cannot be reproduced on the device, so parity with the reference is in distribution (tests check the subset /
range / uniformity properties); THIS file pins the device kernel's integer arithmetic bit for bit (pure uint32 numpy).
Only tests/ may import it.
"""
import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def mix32(x):
    x = np.uint64(x) & M32
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x7feb352d)) & M32
    x ^= x >> np.uint64(15); x = (x * np.uint64(0x846ca68b)) & M32
    x ^= x >> np.uint64(16)
    return x & M32


def bag_key(seed, b):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    lo, hi = seed & 0xFFFFFFFF, seed >> 32
    inner = mix32((hi + 0x9E3779B9 * (b + 1)) & 0xFFFFFFFF)
    return int(mix32(np.uint64(lo) ^ inner))


def feistel_perm(i, n, key):
    if n <= 1:
        return 0
    bits = int(n - 1).bit_length()
    h = (bits + 1) >> 1
    mask = (1 << h) - 1
    x = i
    while True:
        l, r = x >> h, x & mask
        for rnd in range(4):
            f = int(mix32((r * 0x9E3779B1 + key + rnd * 0x85EBCA6B) & 0xFFFFFFFF)) & mask
            l, r = r, l ^ f
        x = (l << h) | r
        if x < n:
            return x


def sample_indices(bag_len, n_sample, seed):
    """[n_bags, n_sample] int32 rows chosen by mdl_sample_gather_f32 (-1 where the bag is missing)."""
    out = np.full((len(bag_len), n_sample), -1, dtype=np.int32)
    for b, n in enumerate(bag_len):
        n = int(n)
        if n <= 0:
            continue
        key = bag_key(seed, b)
        for s in range(n_sample):
            if n >= n_sample:
                out[b, s] = feistel_perm(s, n, key)
            else:
                out[b, s] = int(mix32(np.uint64(key) ^ mix32((s + 0x632BE5AB) & 0xFFFFFFFF))) % n
    return out
