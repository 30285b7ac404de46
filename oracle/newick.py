"""oracle/newick.py -- TEST INFRASTRUCTURE ONLY.

Pure-Python restatement of the reference's extended-Newick reader (io/newick.hpp:100-272) used by the tests
to turn Newick text into arrays for the oracle WITHOUT going through the product's C++ reader.
Follows the reference routine by routine:
  read_tree      newick.hpp:125-135   skip white space, expect ';', read one subtree
  read_subtree   newick.hpp:159-190   id = nodes seen so far (before descending), name, hybrid lookup
  read_internal  newick.hpp:193-203   ')' branchset '('
  read_branchset newick.hpp:206-224   comma separated, duplicate arc = error
  read_branch    newick.hpp:228-233   length, subtree, THEN append (parent, child)
  get_hybrid_info newick.hpp:236-248  text after the last '#', first digit onwards = key
  read_length    newick.hpp:251-259, read_name newick.hpp:263-271
Pinned against the reference itself: tests/golden/newick_ref.json (made by oracle/make_golden.py from
oracle/_ref/ref_tc parse).
"""
import sys


class MalformedNewick(Exception):
    pass


def parse_newick(s):
    """returns (num_nodes, edges [(tail, head)...] in reference order, labels {node: name})"""
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 20000))
    st = {"back": len(s) - 1, "n": 0}
    edges, labels, hybrids = [], {}, {}
    LEN_CHARS = set("9876543210.-+Ee")

    def skip_ws():
        while st["back"] >= 0 and s[st["back"]].isspace():
            st["back"] -= 1

    def read_name():
        b = st["back"]
        i = b
        while i >= 0 and s[i] not in "(),":
            i -= 1
        st["back"] = i
        return s[i + 1:b + 1]

    def read_length():
        i = st["back"]
        while i >= 0 and s[i] in LEN_CHARS:
            i -= 1
        if i >= 0 and s[i] == ":":
            st["back"] = i - 1

    def read_subtree():
        skip_ws()
        name = read_name()
        sharp = name.rfind("#")
        if sharp != -1:
            digits = [k for k in range(sharp, len(name)) if name[k].isdigit()]
            if not digits:
                raise MalformedNewick("found '#' but no hybrid number: %r" % name)
            k = digits[0]
            j = k
            while j < len(name) and name[j].isdigit():
                j += 1
            key = int(name[k:j])
            if key in hybrids:
                root = hybrids[key]
            else:
                root = st["n"]
                st["n"] += 1
                hybrids[key] = root
                if name[:sharp]:
                    labels[root] = name[:sharp]
        else:
            root = st["n"]
            st["n"] += 1
            if name:
                labels[root] = name
        if st["back"] > 0 and s[st["back"]] == ")":
            st["back"] -= 1
            seen = set()
            while True:
                read_length()
                child = read_subtree()
                edges.append((root, child))
                if child in seen:
                    raise MalformedNewick("read double edge %d --> %d" % (root, child))
                seen.add(child)
                if st["back"] < 0:
                    raise MalformedNewick("unmatched ')'")
                if s[st["back"]] == ",":
                    st["back"] -= 1
                    continue
                break
            if s[st["back"]] != "(":
                raise MalformedNewick("expected '(' but got %r" % s[st["back"]])
            st["back"] -= 1
        skip_ws()
        return root

    skip_ws()
    if st["back"] >= 0:
        if s[st["back"]] != ";":
            raise MalformedNewick("expected ';'")
        st["back"] -= 1
        read_subtree()
    return st["n"], edges, labels


def write_newick(n, edges, labels):
    """extended Newick of a rooted DAG given as (n, [(tail, head)], {node: name}); a node with several parents is written out
    below the first parent that reaches it (as `(...)name#Hk`) and referenced as `#Hk` below the others (the format
    io/newick.hpp:29-55 writes and :64-272 reads).  Children are written in the order of `edges`."""
    ch = [[] for _ in range(n)]
    indeg = [0] * n
    for t, h in edges:
        ch[t].append(h); indeg[h] += 1
    roots = [v for v in range(n) if indeg[v] == 0]
    assert len(roots) == 1, "need exactly one root"
    hyb, done = {}, set()
    out = []

    def name(v):
        s = labels.get(v, "") if isinstance(labels, dict) else labels[v]
        if indeg[v] >= 2:
            s += "#H%d" % hyb.setdefault(v, len(hyb) + 1)
        return s
    # iterative DFS
    stack = [(roots[0], 0)]
    while stack:
        v, k = stack.pop()
        if k == 0 and indeg[v] >= 2 and v in done:
            out.append(name(v)); continue
        if k == 0:
            done.add(v)
            if ch[v]:
                out.append("(")
        if k < len(ch[v]):
            if k > 0:
                out.append(",")
            stack.append((v, k + 1)); stack.append((ch[v][k], 0))
        else:
            if ch[v]:
                out.append(")")
            out.append(name(v))
    return "".join(out) + ";"
