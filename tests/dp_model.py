"""CPU restatement of the DEVICE algorithm of the general tree-in-tree DP (csrc/dp_kernels.cuh: dp_node / dp_who_flow_kernel,
dp_perfect_matching), written the way the kernel works -- not the way the oracle does (oracle/pt_oracle.c: orc_mul_who_displays is
definition-based: per host node, a matching over its children) -- so that the kernel's logic can be checked on a machine without a GPU:

  * dataflow over the guest tree: a node is processed when its last child has been (here: a work list fed by per-node counters);
  * lists are host PREORDER POSITIONS; a guest leaf's list = the positions of the host leaves with its label, ascending;
  * a node whose children all have ONE candidate: sorted positions, all consecutive LCAs equal v = LCA(first, last), no candidate is v;
  * otherwise: merge keys (position << CHILD_BITS | child index), one stack sweep over the sorted keys with the LCA of consecutive
    candidates, an induced node is finished when it leaves the stack, child sets are W = ceil(d / 64) 64-bit words per induced node,
    Kuhn's augmenting paths with explicit L / it / chosen arrays, accepted nodes shadow their ancestors (minimality).

Test infrastructure only."""

CHILD_BITS = 20
MASK64 = (1 << 64) - 1


class Host:
    def __init__(self, parent):
        n = len(parent)
        self.parent = list(parent)
        kids = [[] for _ in range(n)]
        root = -1
        for v, p in enumerate(parent):
            if p < 0:
                root = v
            else:
                kids[p].append(v)
        self.kids = kids
        self.root = root
        self.pos = [0] * n          # preorder position of a node
        self.nodeat = [0] * n       # node at a position
        self.depth = [0] * n
        st = [root]
        k = 0
        while st:
            v = st.pop()
            self.pos[v] = k
            self.nodeat[k] = v
            k += 1
            for c in reversed(kids[v]):
                self.depth[c] = self.depth[v] + 1
                st.append(c)

    def lca(self, a, b):
        while a != b:
            if self.depth[a] < self.depth[b]:
                a, b = b, a
            a = self.parent[a]
        return a


def perfect_matching(d, r, W, cov, ridx):
    """dp_perfect_matching: all d left nodes matched?  right y adjacent to left i iff bit i of cov[ridx[y]] (W words)"""
    if r < d:
        return False
    for w in range(W):
        allw = 0
        for y in range(r):
            allw |= cov[ridx[y]][w]
        bits = d - 64 * w
        full = MASK64 if bits >= 64 else (1 << bits) - 1
        if allw & full != full:
            return False
    matchR = [-1] * r
    L, it, chosen = [0] * d, [0] * d, [0] * d
    for i in range(d):
        seen = [0] * r
        depth = 0
        L[0] = i
        it[0] = 0
        ok = False
        while depth >= 0:
            l = L[depth]
            found = -1
            while it[depth] < r:
                y = it[depth]
                it[depth] += 1
                if (cov[ridx[y]][l >> 6] >> (l & 63)) & 1 and not seen[y]:
                    seen[y] = 1
                    found = y
                    break
            if found < 0:
                depth -= 1
                continue
            chosen[depth] = found
            if matchR[found] < 0:
                for k in range(depth + 1):
                    matchR[chosen[k]] = L[k]
                ok = True
                break
            if depth + 1 >= d:
                break
            L[depth + 1] = matchR[found]
            it[depth + 1] = 0
            depth += 1
        if not ok:
            return False
    return True


def node_general(H, lists):
    """the general case of dp_node: lists = the candidate position lists of the d live children; returns M(u) as positions"""
    d = len(lists)
    W = (d + 63) >> 6
    keys = sorted((p << CHILD_BITS) | ci for ci, lst in enumerate(lists) for p in lst)
    vpos, vdepth, vfirst, vnext, vnch, vsb, cov = [], [], [], [], [], [], []
    result = []

    def newnode(p):
        vpos.append(p); vdepth.append(H.depth[H.nodeat[p]]); vfirst.append(-1); vnext.append(-1); vnch.append(0); vsb.append(0)
        cov.append([0] * W)
        return len(vpos) - 1

    def finish(x, q):
        succ = 0
        if not vsb[x] and vnch[x] >= d:
            ridx = []
            y = vfirst[x]
            while y >= 0:
                ridx.append(y)
                y = vnext[y]
            if perfect_matching(d, len(ridx), W, cov, ridx):
                result.append(vpos[x])
                succ = 1
        if q >= 0:
            for w in range(W):
                cov[q][w] |= cov[x][w]
            vsb[q] |= succ | vsb[x]
            vnext[x] = vfirst[q]
            vfirst[q] = x
            vnch[q] += 1

    stack = []
    cur = -1
    for key in keys:
        p, ci = key >> CHILD_BITS, key & ((1 << CHILD_BITS) - 1)
        if cur >= 0 and vpos[cur] == p:
            cov[cur][ci >> 6] |= 1 << (ci & 63)
            continue
        w = newnode(p)
        cov[w][ci >> 6] |= 1 << (ci & 63)
        cur = w
        if not stack:
            stack.append(w)
            continue
        t = stack[-1]
        lnode = H.lca(H.nodeat[vpos[t]], H.nodeat[p])
        lp, ld = H.pos[lnode], H.depth[lnode]
        if lp != vpos[t]:
            while len(stack) >= 2 and vdepth[stack[-2]] >= ld:
                finish(stack[-1], stack[-2])
                stack.pop()
            if vpos[stack[-1]] != lp:
                Lx = newnode(lp)
                finish(stack[-1], Lx)
                stack[-1] = Lx
        stack.append(w)
    while len(stack) >= 2:
        finish(stack[-1], stack[-2])
        stack.pop()
    if len(stack) == 1:
        finish(stack[0], -1)
    return result


def node_single(H, ps):
    """every child has one candidate: closed form over the sorted positions; returns [] or [position of v]"""
    ps = sorted(ps)
    v = H.lca(H.nodeat[ps[0]], H.nodeat[ps[-1]])
    vp = H.pos[v]
    if any(p == vp for p in ps):
        return []
    for a in range(len(ps) - 1):
        if ps[a] == ps[a + 1] or H.lca(H.nodeat[ps[a]], H.nodeat[ps[a + 1]]) != v:
            return []
    return [vp]


def who_lists(host_parent, host_label, guest_parent, guest_label):
    """per guest node the host nodes that minimally display it (node ids, ascending in host preorder), [] if none / cleaned away"""
    H = Host(host_parent)
    nt = len(guest_parent)
    by_label = {}
    for v, l in enumerate(host_label):
        if l >= 0 and not H.kids[v]:
            by_label.setdefault(l, []).append(H.pos[v])
    for l in by_label:
        by_label[l].sort()
    gk = [[] for _ in range(nt)]
    for u, p in enumerate(guest_parent):
        if p >= 0:
            gk[p].append(u)
    pending = [len(k) for k in gk]
    DEAD = None
    M = [None] * nt
    state = [0] * nt      # 0 not computed, 1 computed, 2 dead
    work = [u for u in range(nt) if not gk[u]]
    done = 0
    while work:
        u = work.pop()
        while True:
            if not gk[u]:
                if guest_label[u] < 0:
                    state[u] = 2
                else:
                    M[u] = list(by_label.get(guest_label[u], [])); state[u] = 1
            else:
                live = [c for c in gk[u] if state[c] != 2]
                if not live:
                    state[u] = 2
                elif any(not M[c] for c in live):
                    M[u] = []; state[u] = 1
                elif len(live) == 1:
                    M[u] = M[live[0]]; state[u] = 1
                elif all(len(M[c]) == 1 for c in live):
                    M[u] = node_single(H, [M[c][0] for c in live]); state[u] = 1
                else:
                    M[u] = node_general(H, [M[c] for c in live]); state[u] = 1
            done += 1
            p = guest_parent[u]
            if p < 0:
                break
            pending[p] -= 1
            if pending[p] != 0:
                break
            u = p
    assert done == nt
    return [([] if state[u] != 1 else [H.nodeat[p] for p in M[u]]) for u in range(nt)]
