"""Oracle (TEST INFRASTRUCTURE): tree flattening rules of the hot path.

Restates the part of the reference's tree layer that decides ordering on the
likelihood path (paths relative to /root/reference/src/Bpp/Phyl/):

* TreeTemplateTools::parenthesisToTree              (TreeTemplateTools.cpp)
* TreeTemplate::unroot                              (TreeTemplate.h:244-284)
* TreeTemplateTools::getNodes  (post-order, root last)   (TreeTemplateTools.h:354-361)
* TreeTemplateTools::getLeaves (pre-order leaf order)    (TreeTemplateTools.h:96-106)
* AbstractHomogeneousTreeLikelihood::init_ / initBranchLengthsParameters
  (Likelihood/AbstractHomogeneousTreeLikelihood.cpp:140-166, 305-337):
  ``BrLen<i>`` <-> i-th post-order node with the root dropped, lengths clamped
  to [1e-6, 1e4].
"""
from __future__ import annotations

import numpy as np

MIN_BRLEN = 1e-6
MAX_BRLEN = 1e4


class Node:
    __slots__ = ("name", "length", "sons", "father", "id")

    def __init__(self, name=None, length=None):
        self.name = name
        self.length = length
        self.sons = []
        self.father = None
        self.id = -1

    def is_leaf(self):
        return not self.sons

    def add_son(self, s):
        self.sons.append(s)
        s.father = self


def parse_newick(text: str) -> Node:
    """Minimal Newick reader (names, branch lengths, nested parentheses)."""
    s = text.strip()
    if s.endswith(";"):
        s = s[:-1]
    pos = 0

    def skip():
        nonlocal pos
        while pos < len(s) and s[pos].isspace():
            pos += 1

    def parse():
        nonlocal pos
        skip()
        n = Node()
        if s[pos] == "(":
            pos += 1
            while True:
                n.add_son(parse())
                skip()
                if s[pos] == ",":
                    pos += 1
                    continue
                if s[pos] == ")":
                    pos += 1
                    break
                raise ValueError("bad newick at %d" % pos)
        skip()
        st = pos
        while pos < len(s) and s[pos] not in ",():;":
            pos += 1
        nm = s[st:pos].strip()
        if nm:
            n.name = nm
        skip()
        if pos < len(s) and s[pos] == ":":
            pos += 1
            st = pos
            while pos < len(s) and s[pos] not in ",();":
                pos += 1
            n.length = float(s[st:pos])
        return n

    root = parse()
    return root


def unroot(root: Node) -> Node:
    """TreeTemplate::unroot (TreeTemplate.h:244-284)."""
    if len(root.sons) != 2:
        raise ValueError("UnrootedTreeException")
    s1, s2 = root.sons
    if s1.is_leaf() and s2.is_leaf():
        return root
    if s1.is_leaf():
        s1, s2 = s2, s1
    if s1.length is not None:
        s2.length = s1.length + s2.length if s2.length is not None else s1.length
        s1.length = None
    s1.father = None
    s1.add_son(s2)
    return s1


def postorder(root: Node):
    """TreeTemplateTools::getNodes: sons first, node last (root is the last entry)."""
    out = []
    stack = [(root, 0)]
    while stack:
        n, i = stack.pop()
        if i < len(n.sons):
            stack.append((n, i + 1))
            stack.append((n.sons[i], 0))
        else:
            out.append(n)
    return out


def leaves(root: Node):
    """TreeTemplateTools::getLeaves: pre-order leaf order."""
    out = []
    stack = [root]
    while stack:
        n = stack.pop()
        if n.is_leaf():
            out.append(n)
        for s in reversed(n.sons):
            stack.append(s)
    return out


class FlatTree:
    """Flattened topology handed to both the oracle loops and the C ABI.

    Node ids are post-order positions (root = n_nodes-1), so ``BrLen<i>`` is the
    branch above node i -- the reference's nodes_[i] (init_, :155-157).
    """

    def __init__(self, root: Node, check_rooted: bool = True):
        if check_rooted and len(root.sons) == 2:
            root = unroot(root)
        self.root_node = root
        nodes = postorder(root)
        for i, n in enumerate(nodes):
            n.id = i
        self.nodes = nodes
        self.n_nodes = len(nodes)
        self.root = self.n_nodes - 1
        self.parent = np.array([n.father.id if n.father is not None else -1 for n in nodes], np.int32)
        self.children = [[s.id for s in n.sons] for n in nodes]
        self.is_leaf = np.array([n.is_leaf() for n in nodes], bool)
        self.leaf_ids = [n.id for n in leaves(root)]            # pre-order leaf order
        self.leaf_names = [nodes[i].name for i in self.leaf_ids]
        bl = np.zeros(self.n_nodes)
        for i, n in enumerate(nodes[:-1]):
            d = MIN_BRLEN if n.length is None else n.length
            bl[i] = min(max(d, MIN_BRLEN), MAX_BRLEN)             # :305-337
        self.brlen = bl                                           # brlen[root] unused

    def csr(self):
        off = np.zeros(self.n_nodes + 1, np.int32)
        flat = []
        for i, ch in enumerate(self.children):
            off[i + 1] = off[i] + len(ch)
            flat.extend(ch)
        return off, np.array(flat, np.int32)


def clock_parameters(flat: FlatTree):
    """RHomogeneousClockTreeLikelihood::initBranchLengthsParameters (Likelihood/RHomogeneousClockTreeLikelihood.cpp:121-157):
    heights by TreeTemplateTools::getHeights (TreeTemplateTools.cpp:173-186: the LONGEST path to a leaf below the node),
    ``TotalHeight`` = height of the root, ``HeightP<id>`` = height / father's height for every internal non-root node.
    Returns (total_height, {node id: height proportion})."""
    h = np.zeros(flat.n_nodes)
    for nid in range(flat.n_nodes):                      # post-order: sons first
        for s in flat.children[nid]:
            h[nid] = max(h[nid], h[s] + flat.brlen[s])
    hp = {nid: h[nid] / h[int(flat.parent[nid])] for nid in range(flat.n_nodes - 1) if not flat.is_leaf[nid]}
    return float(h[flat.root]), hp


def clock_branch_lengths(flat: FlatTree, total_height: float, height_p: dict, min_brlen: float = 0.0):
    """RHomogeneousClockTreeLikelihood::computeBranchLengthsFromHeights (:161-179): a leaf son hangs at the height of its
    father, an internal son at HeightP * father's height; lengths are floored at minimumBrLen_ (0 for the clock class, :87)."""
    bl = np.zeros(flat.n_nodes)
    height = {flat.root: total_height}
    for nid in range(flat.n_nodes - 1, -1, -1):          # fathers before sons
        if nid not in height:
            continue
        for s in flat.children[nid]:
            if flat.is_leaf[s]:
                bl[s] = max(min_brlen, height[nid])
            else:
                height[s] = height_p[s] * height[nid]
                bl[s] = max(min_brlen, height[nid] - height[s])
    return bl


def random_tree(n_taxa: int, rng: np.random.Generator, mean_brlen: float = 0.05, rooted: bool = False) -> Node:
    """Random binary topology by sequential random attachment; Exp(mean) lengths.
    Unrooted: 3-son root (what init_ produces after unroot()); rooted: 2-son root."""
    names = ["t%d" % i for i in range(n_taxa)]

    def L():
        return float(rng.exponential(mean_brlen))

    a, b = Node(names[0], L()), Node(names[1], L())
    root = Node()
    root.add_son(a)
    root.add_son(b)
    edges = [a, b]            # nodes identified with the branch above them
    start = 2
    if not rooted and n_taxa >= 3:
        c = Node(names[2], L())
        root.add_son(c)
        edges.append(c)
        start = 3
    for k in range(start, n_taxa):
        e = edges[int(rng.integers(len(edges)))]
        f = e.father
        mid = Node(None, L())
        idx = f.sons.index(e)
        f.sons[idx] = mid
        mid.father = f
        mid.add_son(e)
        leaf = Node(names[k], L())
        mid.add_son(leaf)
        edges.extend([mid, leaf])
    return root


def to_newick(n: Node) -> str:
    s = ""
    if n.sons:
        s += "(" + ",".join(to_newick(c) for c in n.sons) + ")"
    if n.name:
        s += n.name
    if n.length is not None:
        s += ":%r" % n.length
    return s + (";" if n.father is None else "")
