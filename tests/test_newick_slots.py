"""The idea behind nw_parse_line_warp (csrc/newick_kernels.cu), restated in plain Python and checked on the CPU against the reader
that is pinned to the reference (oracle/newick.py, tests/golden/newick_ref.json): the right-to-left scan of io/newick.hpp:159-272
can be COUNTED instead of walked --
  * one token slot left of the ';' and left of every ')' and ','; slot order = node order (repeated hybrid numbers fold onto
    their first slot),
  * one arc per ',' and '(' in reading order: from the owner of the open group (the nearest ')' to the right that opened this
    depth) to the subtree just finished (the slot opened last, or -- right after a '(' -- the owner that '(' closed),
  * a line takes this path only if it is well formed; everything else belongs to the sequential scan.
This mirrors the kernel's chunked computation with one "chunk" as long as the line; the CUDA code itself is tested on the GPU
(tests/test_gpu_newick.py)."""
import random

import pytest

from conftest import load_golden


def parse_by_counting(s):
    """(num_nodes, arcs, labels) or None when the line has to go the sequential way"""
    WS = " \t\n\r\v\f"
    if not s or s[-1] != ";":
        return None
    n = len(s)
    order = list(range(n - 2, -1, -1))                      # reading order: right to left, after the ';'
    sp, last = 0, ";"
    depth_before, right_delim = {}, {}
    for p in order:
        c = s[p]
        depth_before[p], right_delim[p] = sp, last
        if c == ")":
            if p == 0 or last == "(":
                return None
            sp += 1
        elif c in "(,":
            if sp < 1:
                return None
            if c == "(":
                sp -= 1
        elif last == "(" and c not in WS:
            return None
        if c in "(),":
            last = c
    if sp != 0:
        return None
    # slots and arcs by counting
    slot_start = [n - 2]                                    # position where the slot's token begins (going left)
    owner = {}                                              # depth -> slot of the ')' that opened it (the kernel: match.any, else st[])
    arcs = []
    for p in order:
        c = s[p]
        if c not in "(),":
            continue
        k = depth_before[p]
        latest = len(slot_start) - 1
        if c == ")":
            owner[k] = latest
        else:
            child = owner[k] if right_delim[p] == "(" else latest
            arcs.append((owner[k - 1], child))
        if c in "),":
            slot_start.append(p - 1)
    # tokens, one slot at a time, with the sequential rules
    names, keys = [], []
    for i, b in enumerate(slot_start):
        if i:                                               # ":<number>" only directly right of the name, never for the root
            j = b
            while j >= 0 and s[j] in "0123456789.-+Ee":
                j -= 1
            if j >= 0 and s[j] == ":":
                b = j - 1
        while b >= 0 and s[b] in WS:
            b -= 1
        e = b
        while e >= 0 and s[e] not in "(),":
            e -= 1
        tok = s[e + 1:b + 1]
        sharp = tok.rfind("#")
        if sharp < 0:
            names.append(tok); keys.append(None)
            continue
        d = sharp
        while d < len(tok) and not tok[d].isdigit():
            d += 1
        if d == len(tok):
            return None
        d2 = d
        while d2 < len(tok) and tok[d2].isdigit():
            d2 += 1
        names.append(tok[:sharp]); keys.append(int(tok[d:d2]))
    first, node_of, labels = {}, [], {}
    for i, key in enumerate(keys):
        if key is not None and key in first:
            node_of.append(node_of[first[key]])
            continue
        if key is not None:
            first[key] = i
        node_of.append(len(labels))
        labels[len(labels)] = names[i]
    return len(labels), [(node_of[a], node_of[b]) for a, b in arcs], labels


def _same(s, oracle_newick):
    got = parse_by_counting(s)
    try:
        n, edges, labels = oracle_newick.parse_newick(s)
    except Exception:
        return got is None or "either"                     # malformed for the reader: the counting path may refuse it, nothing to compare
    if got is None:
        return "sequential"
    assert got[0] == n and got[1] == [tuple(e) for e in edges], s
    assert {v: x for v, x in got[2].items() if x != ""} == {v: x for v, x in labels.items() if x != ""}, s
    return "counted"


def test_counting_equals_the_pinned_reader_on_the_reference_strings():
    import oracle.newick as onw
    g = load_golden("newick_ref.json")
    kinds = [_same(it["newick"], onw) for it in g["items"]]
    assert kinds.count("counted") >= len(kinds) // 2


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_counting_equals_the_pinned_reader_on_random_and_mutated_strings(seed):
    import oracle.newick as onw
    rnd = random.Random(seed)

    def tree(k, hyb):
        if k == 1:
            t = rnd.choice(["a", "b7", "12", "3e", "x.y", "", " q"]) + rnd.choice(["", "", ":1.5", ":2e-3", ":"])
            return t if t.strip() or rnd.random() < 0.5 else "z"
        c = rnd.randint(1, k - 1)
        kids = [tree(c, hyb), tree(k - c, hyb)]
        if rnd.random() < 0.2:
            kids.append(tree(1, hyb))
        if hyb and rnd.random() < 0.15:
            kids.append("#H%d" % rnd.choice(hyb))
        return "(" + rnd.choice(["", " "]) + ",".join(kids) + ")" + rnd.choice(["", "i", "n1", "(x)y#H9"[5:]]) + rnd.choice(["", ":0.1", " "])
    counted = 0
    for it in range(400):
        hyb = [1, 2, 33] if it % 2 else []
        s = tree(rnd.randint(1, 25), hyb).rstrip() + ";"
        if it % 4 == 3:                                     # mutate: the counting path must refuse or agree, never differ
            t = list(s)
            for _ in range(rnd.randint(1, 2)):
                k = rnd.randrange(len(t))
                t[k] = rnd.choice("(),;#: a1")
            s = "".join(t)
        counted += _same(s, onw) == "counted"
    assert counted > 150


def _label_ids_by_chunks(lines):
    """nw_label_chunks: the nodes of the lines, in order, 32 at a time; inside a chunk every node claims or joins its label's slot, the
    slot keeps the SMALLEST node index seen in THIS chunk (a claim made by an earlier chunk or line stays), and the nodes that own a
    slot afterwards number the new labels in lane order"""
    slot, ids, count = {}, [], 0                            # slot: label -> region index of its owner
    for names, base in lines:                               # base = where the line's nodes sit in the pair's region: the HOST line comes
                                                            # first even when it is the second line of the pair (larger indices)
        for c0 in range(0, len(names), 32):
            chunk = list(range(c0, min(c0 + 32, len(names))))
            lanes = chunk[:]
            random.Random(c0).shuffle(lanes)                # the hardware order of the claims inside a chunk is arbitrary
            for v in lanes:
                if names[v] == "":
                    continue
                me = base + v
                if names[v] not in slot:
                    slot[names[v]] = me
                elif me < slot[names[v]] < base + chunk[-1] + 1 and slot[names[v]] >= base + c0:
                    slot[names[v]] = me                     # atomicMin, only against an owner of the same chunk
            fresh = [v for v in chunk if names[v] != "" and slot[names[v]] == base + v]
            new_id = {base + v: count + i for i, v in enumerate(fresh)}
            count += len(fresh)
            for v in chunk:
                if names[v] == "":
                    ids.append(-1)
                elif base + v in new_id:
                    ids.append(new_id[base + v]); slot_id[names[v]] = new_id[base + v]
                else:
                    ids.append(slot_id[names[v]])
    return ids


slot_id = {}


@pytest.mark.parametrize("seed", [5, 6])
def test_label_ids_by_chunks_are_first_appearance_ids(seed):
    rnd = random.Random(seed)
    for it in range(300):
        two = [[rnd.choice(["", "a", "b", "c", "d", "e", "f"]) + (str(rnd.randint(0, 30)) if rnd.random() < 0.8 else "") for _ in range(rnd.randint(1, 100))] for _ in range(2)]
        swapped = it % 2 == 1                               # the guest line first: the host (processed first) is the second line
        lines = [(two[1], len(two[0]) + 2), (two[0], 0)] if swapped else [(two[0], 0), (two[1], len(two[0]) + 2)]
        slot_id.clear()
        got = _label_ids_by_chunks(lines)
        seen, want = {}, []
        for names, _ in lines:
            for x in names:
                want.append(-1 if x == "" else seen.setdefault(x, len(seen)))
        assert got == want, lines


def _csr_by_chunks(n, arcs):
    """nw_warp_csr: arcs 32 at a time in arc order; lanes sharing a tail rank themselves (match.any), the lowest lane moves the cursor"""
    cnt = [0] * n
    for t, _ in arcs:
        cnt[t] += 1
    off = [0]
    for v in range(n):
        off.append(off[-1] + cnt[v])
    cur, adj = off[1:], [None] * len(arcs)
    for e0 in range(0, len(arcs), 32):
        chunk = arcs[e0:e0 + 32]
        snapshot = cur[:]                                   # every lane reads its cursor before any lane writes
        for lane, (t, h) in enumerate(chunk):
            rank = sum(1 for tt, _ in chunk[:lane] if tt == t)
            adj[snapshot[t] - 1 - rank] = h
        for t in {t for t, _ in chunk}:
            cur[t] = snapshot[t] - sum(1 for tt, _ in chunk if tt == t)
    return off, adj


def test_csr_by_chunks_keeps_the_sequential_child_order():
    rnd = random.Random(9)
    for it in range(300):
        n = rnd.randint(1, 80)
        arcs = [(rnd.randrange(n) if rnd.random() < 0.7 else 0, rnd.randrange(n)) for _ in range(rnd.randint(0, 200))]
        off, adj = _csr_by_chunks(n, arcs)
        cur, want = off[1:], [None] * len(arcs)
        for t, h in arcs:                                   # Phylogeny::build order: A[--cursor[tail]] = head over ascending arcs
            cur[t] -= 1
            want[cur[t]] = h
        assert adj == want
