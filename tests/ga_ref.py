"""numpy restatement of the device GA (csrc/prl_ga.cu) -- TEST INFRASTRUCTURE. Same Philox4x32-10
counters, same fixed-point weights, same inverse-CDF picks, so the children of a generation can be
compared with the GPU's bit for bit. Semantics follow the reference's src/main.rs:72-160."""
import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF
DRAW_PARENTS = 0x40000000


def philox4x32(seed, index, generation, draw):
    c = [index & MASK, (index >> 32) & MASK, generation & MASK, draw & MASK]
    k0, k1 = seed & MASK, (seed >> 32) & MASK
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c[3] ^ k1) & MASK, p0 & MASK]
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c


def init_genomes(n_agents, genome_len, seed):
    """main.rs:72-82: three fair coin flips per gene."""
    out = np.zeros((n_agents, genome_len), np.uint8)
    for a in range(n_agents):
        for g in range(genome_len):
            r = philox4x32(seed, a, 0, g)
            out[a, g] = (r[0] >> 31) | ((r[1] >> 31) << 1) | ((r[2] >> 31) << 2)
    return out


def best(scores):
    """main.rs:87-111: (first minimum's index, min, max)."""
    i = int(np.argmin(scores))  # numpy returns the first minimum
    return i, np.float32(scores[i]), np.float32(scores.max())


def fixed_weights(scores):
    mx = np.float32(scores.max())
    w = (mx + np.float32(1.0)) - scores.astype(np.float32)   # f32, like `max_score + 1.0 - score`
    w = np.where(w > 0, w, 0).astype(np.float64)
    return [int(np.rint(x * 1048576.0)) for x in w]


def pick(cdf, total, hi, lo):
    target = ((((hi << 32) | lo) * total) >> 64)
    a, b = 0, len(cdf) - 1
    while a < b:
        m = (a + b) >> 1
        if cdf[m] > target:
            b = m
        else:
            a = m + 1
    return a


def breed(parents, scores, generation, seed, mutation_rate, keep_best, children=None):
    """main.rs:126-158. parents: (N, L) uint8, scores: (N,) f32 -> children (N, L) uint8 plus the
    chosen parent indices. `children`: iterable of child indices to produce (default all)."""
    n, L = parents.shape
    w = fixed_weights(scores)
    cdf = list(np.cumsum(np.array(w, dtype=object)))
    total = cdf[-1]
    bi, _, _ = best(scores)
    rate = np.float32(mutation_rate)
    out, chosen = {}, {}
    for child in (range(n) if children is None else children):
        if keep_best and child == 0:
            out[child] = parents[bi].copy()
            chosen[child] = (bi, bi)
            continue
        r = philox4x32(seed, child, generation, DRAW_PARENTS)
        p1, p2 = pick(cdf, total, r[0], r[1]), pick(cdf, total, r[2], r[3])
        attempt = 1
        while p2 == p1 and n > 1 and attempt < 64:
            r = philox4x32(seed, child, generation, DRAW_PARENTS + attempt)
            p2 = pick(cdf, total, r[0], r[1])
            if p2 == p1:
                p2 = pick(cdf, total, r[2], r[3])
            attempt += 1
        if p2 == p1 and n > 1:
            p2 = (p1 + 1) % n
        g = np.zeros(L, np.uint8)
        for i in range(L):
            r = philox4x32(seed, child, generation, i)
            mv = int(parents[p1, i] if (r[0] >> 31) else parents[p2, i])
            for bit in range(3):
                u = np.float32(r[1 + bit] >> 8) * np.float32(1.0 / 16777216.0)
                if u < rate:
                    mv = (mv & ~(1 << bit)) | ((r[1 + bit] & 1) << bit)
            g[i] = mv & 7
        out[child] = g
        chosen[child] = (p1, p2)
    return out, chosen
