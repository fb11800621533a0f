"""Shared cases for the compilation-cache bloom pre-filter (flare SaltedBloomFilter)."""
import numpy as np


def tu_keys(n=6124, seed=46):
    """Cache-entry keys as GetCxxCacheEntryKey builds them: "yadcc-cxx2-entry-" + 64 hex
    (yadcc/daemon/cache_format.cc:56-64).  6124 = the LLVM-11 target count (README.md:88)."""
    rng = np.random.default_rng(seed)
    return ["yadcc-cxx2-entry-" + rng.bytes(32).hex() for _ in range(n)]


def run_bloom_suite(d, seed=0):
    """Everything observable: filter bytes after Add, lookups (incl. false positives), odd
    key lengths, tiny and yadcc-sized geometries, loading an existing filter."""
    rng = np.random.default_rng(seed)
    out = []
    keys = tu_keys(6124, 46 + seed)
    d.bloom_reset()  # 27 584 639 -> 2^25 bits, 10 hashes
    known = [k for k, m in zip(keys, rng.random(len(keys)) < 0.3) if m]
    d.bloom_add(known)
    trace = [keys[i % len(keys)] for i in range(20000)]
    out.append(d.bloom_possibly_contains(trace).astype(np.uint8))
    b = d.bloom_bytes()
    out.append(np.frombuffer(b, dtype=np.uint8).copy())
    # a small, crowded filter: many false positives
    d.bloom_reset(4096, 3)
    d.bloom_add(keys[:700])
    out.append(d.bloom_possibly_contains(keys).astype(np.uint8))
    out.append(np.frombuffer(d.bloom_bytes(), dtype=np.uint8).copy())
    # key lengths around the XXH64 stripe / tail boundaries
    for ln in (0, 1, 3, 4, 5, 7, 8, 11, 12, 27, 28, 29, 31, 32, 35, 36, 59, 60, 61, 63, 64, 91, 92, 100, 128, 200, 252):
        ks = rng.integers(0, 256, (37, ln), dtype=np.uint8) if ln else np.zeros((1, 0), dtype=np.uint8)
        d.bloom_reset(1 << 14, 7)
        d.bloom_add(ks[: max(1, len(ks) // 2)])
        out.append(np.frombuffer(d.bloom_bytes(), dtype=np.uint8).copy())
        out.append(d.bloom_possibly_contains(ks).astype(np.uint8))
    # tiny geometries (the constructor clamps to >= 8 bits)
    for bits in (1, 8, 9, 16, 31, 64):
        d.bloom_reset(bits, 4)
        d.bloom_add(keys[:3])
        out.append(np.frombuffer(d.bloom_bytes(), dtype=np.uint8).copy())
    # import a filter produced elsewhere
    d.bloom_load(b, 10)
    out.append(d.bloom_possibly_contains(keys[:500]).astype(np.uint8))
    return out
