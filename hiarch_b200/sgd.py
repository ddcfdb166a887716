"""ScaledGenomicDistance (host side, exact integer tables) and the O/E pooling tables.

Mirror of `ScaledGenomicDistance` (publish/new_hic_class.py:788-1003): log2-spaced distance
bins, `size(d) = floor((log2(d*bin_size) + k) / k)`, re-indexed to 0..S, with the maximum
distance trimmed to a power of 2^k (`auto_change_end`). The table is data independent, so it is
computed once on the host with the same float64 expressions as the reference (bit-exact ints)
and shipped to the GPU as two int32 arrays per chromosome (see `oe_tables`).
"""
import json

import numpy as np


class ScaledGenomicDistance:
    """Same constructor and lookups as the reference class (new_hic_class.py:788-875)."""

    def __init__(self, k=.5, s=2e4, e=6e7, bin_size=2e4, auto_change_end=True, input_dict=None):
        self.paras = {'k': k, 's': s, 'e': e, 'bin_size': bin_size,
                      'auto_change_end': auto_change_end}
        if isinstance(input_dict, dict):
            self.paras.update(input_dict)
        elif isinstance(input_dict, str):
            self.paras.update(json.loads(input_dict.split('#sgd: ')[1]))
        elif input_dict is not None:
            raise ValueError(f'Unrecognized type of input_dict: {type(input_dict)}')
        self.size_dict = None
        self.max_dis = None
        self.min_dis = None
        self.max_size = None
        self.n_bins_of_sizes = None
        self.size_ranges = None
        self.get_sgd()

    def get_sgd(self):
        """new_hic_class.py:899-1000."""
        self.auto_change_end = self.paras['auto_change_end']
        self.e, self.s = self.paras['e'], self.paras['s']
        self.k, self.bin_size = self.paras['k'], self.paras['bin_size']
        if self.auto_change_end:
            cur_max_size = (np.log2(self.e) + self.k) / self.k
            e = np.power(2, self.k * np.floor(cur_max_size) - self.k)
        else:
            e = self.e
        self.min_dis = self.s / self.bin_size
        self.max_dis = e / self.bin_size
        size_max = int(np.floor((np.log2(e) + self.k) / self.k)) + 1
        size_dict = {}
        for diff in range(int(self.s), int(e) + 1, int(self.bin_size)):
            size = np.floor((np.log2(diff) + self.k) / self.k)
            if size >= size_max:
                break
            size_dict[int(diff / self.bin_size)] = size
        raw_sizes = sorted(set(size_dict.values()))
        reordered = {raw: i for i, raw in enumerate(raw_sizes)}
        self.size_dict = {d: reordered[v] for d, v in size_dict.items()}
        values = np.array(list(self.size_dict.values()))
        keys = np.array(list(self.size_dict.keys()))
        self.max_size = np.max(values)
        self.n_bins_of_sizes = {s_: int(np.count_nonzero(values == s_)) * 2
                                for s_ in range(len(raw_sizes))}
        self.size_ranges = {s_: [np.min(keys[values == s_]), np.max(keys[values == s_])]
                            for s_ in range(len(raw_sizes))}
        return size_dict

    def __getitem__(self, dis, force_mode=None):
        """Scalar lookup incl. the reference's `dis >= bin_size` quirk (:865-875)."""
        if isinstance(dis, np.ndarray):
            raise NotImplementedError("array lookups are not on the hot path")
        if force_mode is None:
            if not dis < self.bin_size:
                dis = dis / self.bin_size
        elif force_mode == 'raw':
            dis = dis / self.bin_size
        return self.size_dict.get(dis)

    def out_paras_to_string(self):
        return f'#sgd: {json.dumps(self.paras)}\n'


def oe_tables(bin_size, n, k=.2):
    """Data-independent pooling tables of `_sgd_mean` (new_hic_class.py:1182-1216).

    Walks the diagonals exactly like the reference: diagonals whose SGD is None are skipped,
    consecutive ones with the same SGD are pooled, a pooled mean is *stored only when the SGD
    changes* (the last group is never stored), a later store to the same key overwrites.
    Returns (pool_gid[n], use_code[n]) int32 -- see include/hiarch_b200.h (hia_oe_finalize).
    """
    pool_gid = np.full(n, -1, dtype=np.int32)
    use_code = np.full(n, -1, dtype=np.int32)
    if n <= 1:
        return pool_gid, use_code
    sgd = ScaledGenomicDistance(s=bin_size, e=bin_size * n, bin_size=bin_size, k=k)
    look = [sgd[d] for d in range(n)]
    stores = []                       # (key, members) in store order
    cur, members = None, []
    for d in range(n):
        g = look[d]
        if g is None:
            continue
        if cur is None or cur == g:
            members.append(d)
            cur = g
        else:
            stores.append((cur, members))
            members, cur = [d], g
    if not stores:
        # np.max(list({}.keys())) in the reference (:1211)
        raise ValueError("zero-size array to reduction operation maximum which has no identity "
                         f"(chromosome of {n} bins has a single SGD group)")
    last_store = {}
    for idx, (key, mem) in enumerate(stores):
        last_store[key] = idx
        pool_gid[mem] = idx
    fallback = last_store[max(last_store.keys())]
    for d in range(n):
        g = look[d]
        if g is None or g not in last_store:
            use_code[d] = -1 if d == 0 else 2 * fallback
        else:
            use_code[d] = 2 * last_store[g] + 1
    if len(stores) > 512:
        raise ValueError("more than 512 SGD groups")
    return pool_gid, use_code
