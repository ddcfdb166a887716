"""a12: centre-anchor search on a 1-D profile (host side, sequential, integer outputs bit-exact).

Port of `find_anchors_by_inter_rowsum` and friends
(publish/confor/src/S1_preprocess/S1_2_find_anchors/S2_1_find_by_inter_rowsum.py:66-367),
the intra-x variant (S2_2_find_by_intra_x.py:141-184), `change_anchor` / `remove_all_used`
(U3_change_anchor.py:5-33) and the `Anchors` table glue (publish/iutils/read_matrix.py:195-342).
The profile (length n per chromosome) comes from the GPU (`ops.inter_rowsum_profile`, ...);
this logic is O(n * search_len) data-dependent branching on a tiny array and stays on the host.
Regions are plain dicts kept in lists in exactly the order the reference's DataFrame rows have,
because `remove_small_anchors` and `mark_*` depend on that order.
"""
import math

import numpy as np

NAN = float("nan")

MIN_FOR_LARGER_DIFF = .03      # S2_1:181
MIN_FOR_SMALLER_DIFF = .01     # S2_1:182
DECREASE_PER = 1 / 3           # S2_1:290


def get_kernal_len(n, kp, mkl, only_odd=False):
    """S2_1:15-20."""
    kl = int(n * kp)
    kl = mkl if kl < mkl else kl
    if only_odd:
        kl = (kl // 2) * 2 + 1
    return kl


def get_anchor_regions(a, search_per=.1, min_search_len=5):
    """S2_1:66-146. Returns regions (list of dicts with rs, re, cen) sorted by cen."""
    n = len(a)
    search_len = get_kernal_len(n, search_per, min_search_len)
    nhp = [0] * n
    for ip in range(n):                                   # nearest strictly higher point
        found = False
        for step in range(search_len):
            if ip - step >= 0 and a[ip - step] > a[ip]:
                nhp[ip] = ip - step
                found = True
                break
            if ip + step <= n - 1 and a[ip + step] > a[ip]:
                nhp[ip] = ip + step
                found = True
                break
        if not found:
            nhp[ip] = ip
    assign = [0] * n
    regions = {}                                          # cluster id -> dict, insertion = id order
    cluster = 1
    for ip in range(n):
        if nhp[ip] == ip and a[ip] > 0:
            assign[ip] = cluster
            regions[cluster] = {"rs": ip, "re": ip, "cen": ip}
            cluster += 1
    new_join = 1
    while new_join != 0:
        new_join = 0
        for ip in range(n):
            if assign[ip] != 0 or nhp[ip] == ip:
                continue
            pc = nhp[ip]
            if assign[pc] != 0:
                new_join += 1
                inside = [r for r in regions.values() if ip > r["rs"] and ip < r["re"]]
                if inside:
                    assign[ip] = assign[inside[0]["cen"]]
                else:
                    cid = assign[pc]
                    assign[ip] = cid
                    if ip < regions[cid]["rs"]:
                        regions[cid]["rs"] = ip
                    if ip > regions[cid]["re"]:
                        regions[cid]["re"] = ip
    out = sorted(regions.values(), key=lambda r: r["cen"])           # sort_values(by='cen')
    return out


def mark_anchor_state(a, region):
    """S2_1:149-178 for one region (in place)."""
    rs, re, cen = int(region["rs"]), int(region["re"]), int(region["cen"])
    cr = a[cen]
    region["cen_rowsum"] = cr
    before = a[rs:cen]
    if len(before) > 0:
        region["before_min"] = np.min(before)
        region["before_diff"] = cr - np.min(before)
    else:
        region.setdefault("before_min", NAN)
        region.setdefault("before_diff", NAN)
    after = a[cen + 1: re + 1]
    if len(after) > 0:
        region["after_min"] = np.min(after)
        region["after_diff"] = cr - np.min(after)
    else:
        region.setdefault("after_min", NAN)
        region.setdefault("after_diff", NAN)
    return region


def _py_min_key(keys, values):
    """Python's min(dict, key=dict.get) semantics (NaN never compares smaller)."""
    best = 0
    for i in range(1, len(keys)):
        if values[i] < values[best]:
            best = i
    return keys[best]


def remove_small_anchors(regions, a):
    """S2_1:183-287. `regions`: list in DataFrame row order; returns the new list."""
    n = len(a)

    def merge_region(pos):
        region = regions[pos]
        cand_pos, cand_val = [], []
        if pos - 1 >= 0:
            cand_pos.append(pos - 1)
            cand_val.append(region["before_diff"])
        if pos + 1 < len(regions):
            cand_pos.append(pos + 1)
            cand_val.append(region["after_diff"])
        if not cand_pos:
            return [r for i, r in enumerate(regions) if i != pos]
        mpos = _py_min_key(cand_pos, cand_val)
        merged = dict(regions[mpos])
        new_rs = min(merged["rs"], region["rs"])
        new_re = max(merged["re"], region["re"])
        new_cen = merged["cen"] if merged["cen_rowsum"] > region["cen_rowsum"] else int(region["cen"])
        merged["rs"], merged["re"], merged["cen"] = new_rs, new_re, new_cen
        # mark_anchor_state(update_idxes=[merge_idx]) overwrites only when the slice is non-empty;
        # otherwise the merged row keeps its previous value (the reference does not reset to NaN)
        mark_anchor_state(a, merged)
        out = []
        for i, r in enumerate(regions):
            if i == pos:
                continue
            out.append(merged if i == mpos else r)
        return out

    changed = True
    while changed:
        changed = False
        for pos, r in enumerate(regions):
            diffs = []
            if r["rs"] != 0:
                diffs.append(r["before_diff"])
            if r["re"] != n - 1:
                diffs.append(r["after_diff"])
            diffs = [d for d in diffs if not math.isnan(d)]
            if r["rs"] == 0 and r["re"] == n - 1:
                continue
            if len(diffs) == 0 or max(diffs) < MIN_FOR_LARGER_DIFF or min(diffs) < MIN_FOR_SMALLER_DIFF:
                changed = True
                regions = merge_region(pos)
                break
    return regions


def mark_telo_anchors(regions):
    """S2_1:291-318."""
    for r in regions:
        r["type"] = "normal"
    if not regions:
        return regions
    regions = sorted(regions, key=lambda r: r["rs"])                 # sort_values('rs')
    for which in (0, -1):
        r = regions[which]
        before, after = r["before_diff"], r["after_diff"]
        if math.isnan(before) or math.isnan(after):
            r["type"] = "telo"
            continue
        if which == 0 and before < after * DECREASE_PER:
            r["type"] = "telo"
        if which == -1 and after < before * DECREASE_PER:
            r["type"] = "telo"
    return regions


def mark_best_regions(regions, col="cen_rowsum"):
    """S2_1:322-329: the non-telo region with the largest centre value becomes 'used'."""
    cand = [r for r in regions if r["type"] != "telo"]
    if not cand:
        return regions
    best = cand[0]
    for r in cand[1:]:
        if r[col] > best[col]:
            best = r
    best["type"] = "used"
    return regions


def reindex_n_add_cs_ce(regions, chr_start_idx, chr_end_idx):
    """S2_1:333-349."""
    for r in regions:
        for col in ("cen", "rs", "re"):
            r[col] = int(r[col]) + chr_start_idx
        r["cs"] = r["cen"]
        r["ce"] = r["cen"] + 1
        if r["ce"] > chr_end_idx:
            r["cs"] = r["cen"] - 1
            r["ce"] = r["cen"]
    return regions


COLUMNS = ["rs", "re", "cen", "before_min", "before_diff", "after_min", "after_diff", "cen_rowsum",
           "type", "cs", "ce", "chr", "start", "end"]
COLUMNS_X = ["rs", "re", "cen", "cen_intrax", "type", "cs", "ce", "chr", "start", "end"]


class AnchorTable:
    """The rows of `<output>.anchors.txt` (read_matrix.Anchors.loc_df)."""

    def __init__(self, rows, columns):
        self.rows = rows
        self.columns = columns

    def column(self, name):
        return [r[name] for r in self.rows]

    def used_center(self, chro, chr_start=None):
        """Anchors.get_used_chro_loc(chro, chr_idx=True): first 'used' row of the chromosome."""
        for r in self.rows:
            if r["chr"] == chro and r["type"] == "used":
                return r["cen"] - (chr_start or 0)
        return None

    def to_text(self):
        """Same text as `loc_df.to_csv(sep='\\t', index=False)` (S1_2_find_anchor_main.py:41-42)."""
        def fmt(v):
            if isinstance(v, str):
                return v
            if isinstance(v, (int, np.integer)):
                return str(int(v))
            if v != v:
                return ""
            return repr(float(v))
        lines = ["\t".join(self.columns)]
        for r in self.rows:
            lines.append("\t".join(fmt(r[c]) for c in self.columns))
        return "\n".join(lines) + "\n"

    def write(self, path):
        with open(path, "w") as fh:
            fh.write(self.to_text())


def _add_positions(rows, names, seps, bin_starts, bin_ends):
    """Anchors.find_anchor_gen_pos (read_matrix.py:251-272): chr / start / end of the `cs` bin."""
    bounds = np.array([s for s, _ in seps])
    for r in rows:
        c1 = int(np.searchsorted(bounds, r["cs"], side="right") - 1)
        c2 = int(np.searchsorted(bounds, r["ce"], side="right") - 1)
        if c1 != c2:
            raise ValueError("Position of same chromosome are not same.")
        r["chr"] = names[c1]
        r["start"] = float(bin_starts[r["cs"]])
        r["end"] = float(bin_ends[r["cs"]])
    return rows


def find_anchors_by_rowsum(profile, names, seps, bin_starts, bin_ends):
    """find_anchors_by_inter_rowsum (S2_1:352-367) on a smoothed genome-wide profile."""
    profile = np.asarray(profile, dtype=np.float64)
    rows = []
    for (s, e) in seps:
        a = profile[s:e + 1]
        regions = get_anchor_regions(a)
        for r in regions:
            r.update(before_min=NAN, before_diff=NAN, after_min=NAN, after_diff=NAN, cen_rowsum=NAN)
            mark_anchor_state(a, r)
        regions = remove_small_anchors(regions, a)
        regions = mark_telo_anchors(regions)
        regions = mark_best_regions(regions)
        rows += reindex_n_add_cs_ce(regions, s, e)
    return AnchorTable(_add_positions(rows, names, seps, bin_starts, bin_ends), COLUMNS)


def find_anchors_by_intra_x(profile, names, seps, bin_starts, bin_ends):
    """find_anchors_by_intra_x (S2_2:168-184)."""
    profile = np.asarray(profile, dtype=np.float64)
    rows = []
    for (s, e) in seps:
        a = profile[s:e + 1].copy()
        a[a < 0] = 0
        regions = get_anchor_regions(a, .05)
        for r in regions:
            r["cen_intrax"] = a[int(r["cen"])]
            r["type"] = "normal"
        if not regions:
            raise IndexError("index 0 is out of bounds for axis 0 with size 0")   # S2_2:163
        best = regions[0]
        for r in regions[1:]:
            if r["cen_intrax"] > best["cen_intrax"]:
                best = r
        # sort_values(best_value_col, ascending=False): rows come out ordered by cen_intrax
        regions = sorted(regions, key=lambda r: -r["cen_intrax"])
        best["type"] = "used"
        rows += reindex_n_add_cs_ce(regions, s, e)
    return AnchorTable(_add_positions(rows, names, seps, bin_starts, bin_ends), COLUMNS_X)


def remove_all_used(table):
    """U3_change_anchor.py:27-33 (the pipeline's default `no_anchor`)."""
    for r in table.rows:
        if r["type"] == "used":
            r["type"] = "normal"
    return table


def change_anchor(table, names, changes):
    """U3_change_anchor.py:5-24. `changes`: [(chromosome number 1-based, anchor number 1-based or 0)]."""
    for chr_no, anchor_no in changes:
        chro = names[chr_no - 1]
        idxs = [i for i, r in enumerate(table.rows) if r["chr"] == chro]
        idxs.sort(key=lambda i: table.rows[i]["start"])
        for i in idxs:
            if table.rows[i]["type"] == "used":
                table.rows[i]["type"] = "normal"
                break
        if anchor_no != 0:
            table.rows[idxs[anchor_no - 1]]["type"] = "used"
    return table


def parse_anchor_change(text):
    """'1:2,3:1' -> [(1, 2), (3, 1)]  (S1_2_find_anchor_main.py:26-31)."""
    if text is None or text == "":
        return None
    return [(int(p.split(":")[0]), int(p.split(":")[1])) for p in text.split(",")]
