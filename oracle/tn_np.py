"""Plain-Python restatement of the temporal-network alignment (`tn`) of the reference's vendored VCSL code.

TEST INFRASTRUCTURE (see oracle/__init__.py): the checker for csrc/tn_align.cu / localization.TNLocalization,
never a fallback for them.

Restates VSC22-Descriptor-Track-1st/infer/vcsl/vta.py:244-363 (`tn`) including the behaviour of its two library
calls, so that it runs without networkx:
* the graph (vta.py:259-322): one node per (query frame, one of its `top` best reference frames); an edge
  (q_i, r) -> (q_j, r') for q_i < q_j < q_i + tn_max_step when 0 < r' - r < tn_max_step (C2), no reference frame already
  linked from q_i's row lies strictly between (C3, evaluated against the references linked by EARLIER q_j of the same
  q_i) and sims[q_j, r'] >= min_sim (C4); weight = sims[q_j, r'].  The "sink" is the LAST top-k node (vta.py:317-322 --
  node_num - 1 is not a separate node): every node within tn_max_step of it in both axes gets a zero-weight edge to it,
  overwriting the weight of an existing edge.
* networkx 3.x `dag_longest_path` (networkx/algorithms/dag.py): dist[v] = max over predecessors IN EDGE-INSERTION ORDER
  (first maximum wins) of dist[u] + w, or (0, v) without predecessors; the end node is the first maximum of dist in
  topological-sort order.  Predecessor order matters because every edge into a node carries the same weight, so the
  starts of chains tie all the time.  Insertion order of the edges into (q_j, r'): by (q_i ascending, column of r in
  q_i's top-k ascending) -- the np.where order of vta.py:305 -- then the sink links in node order.
  `nx.topological_sort` is restated as Kahn's algorithm by generations with nodes / successors in insertion order.
* the path loop (vta.py:331-362): at most max_path + 1 paths; path edges are re-weighted to 0; source and "sink" nodes
  are dropped from the path; box = (min q, min r, max q, max r); accepted when score / ave_length > min_sim,
  min extent > min_length and IoU with every accepted box < max_iou.

Pinned by tests/test_oracle_tn.py against the reference function itself (vcsl.vta.tn over networkx) on random and
planted similarity matrices.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def topk_rows(sims: np.ndarray, tn_top_k: int) -> Tuple[np.ndarray, np.ndarray]:
    """vta.py:262-265."""
    top = min(tn_top_k, sims.shape[1])
    idx = np.argsort(-sims)[:, :top]
    return idx, np.take_along_axis(sims, idx, axis=-1)


def _iou_max(box, boxes) -> float:
    """max of vta.py:80-95 `iou(box[None], boxes)`; 0 for an empty list."""
    if not boxes:
        return 0.0
    b = np.asarray(box, dtype=np.int64)[None]
    g = np.asarray(boxes, dtype=np.int64)
    lt = np.maximum(b[:, None, :2], g[:, :2])
    rb = np.minimum(b[:, None, 2:], g[:, 2:])
    wh = np.maximum(rb - lt + 1, 0)
    inter = wh[:, :, 0] * wh[:, :, 1]
    ba = (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
    ga = (g[:, 2] - g[:, 0] + 1) * (g[:, 3] - g[:, 1] + 1)
    return float((inter / (ba[:, None] + ga - inter)).max())


def tn_from_topk(idx: np.ndarray, val: np.ndarray, tn_max_step=10, max_path=10, min_sim=0.2, min_length=5,
                 max_iou=0.3) -> List[List[int]]:
    """`tn` given the per-row top-k (reference frame indices, similarities); returns [[q_min, r_min, q_max, r_max]]."""
    Q, top = idx.shape
    n_nodes = 1 + Q * top
    pair = [(-1, -1)] + [(q, int(idx[q, k])) for q in range(Q) for k in range(top)]
    nid = lambda q, k: 1 + q * top + k
    preds: List[List[int]] = [[] for _ in range(n_nodes)]       # predecessor ids in insertion order
    succs: List[List[int]] = [[] for _ in range(n_nodes)]
    weight = {}

    def add_edge(u, v, w):
        if (u, v) not in weight:
            preds[v].append(u)
            succs[u].append(v)
        weight[(u, v)] = w

    for q_i in range(Q):
        r_i = idx[q_i]
        inter: set = set()
        for q_j in range(q_i + 1, min(Q, q_i + tn_max_step)):
            r_j, s_j = idx[q_j], val[q_j]
            linked = []
            for r in range(top):                 # np.where order: rows (destination) first, then columns (source)
                for c in range(top):
                    d = int(r_j[r]) - int(r_i[c])
                    if not (0 < d < tn_max_step):
                        continue
                    if any(int(r_i[c]) < t < int(r_j[r]) for t in inter):
                        continue
                    if not (s_j[r] >= min_sim):
                        continue
                    add_edge(nid(q_i, c), nid(q_j, r), s_j[r])
                    linked.append(int(r_j[r]))
            inter.update(linked)
    last = n_nodes - 1
    pj = pair[last]
    for i in range(last):
        pi = pair[i]
        if pj[0] > pi[0] and pj[1] > pi[1] and pj[0] - pi[0] <= tn_max_step and pj[1] - pi[1] <= tn_max_step:
            add_edge(i, last, 0)

    def topo_order():
        indeg = [len(p) for p in preds]
        gen = [v for v in range(n_nodes) if indeg[v] == 0]
        order = []
        while gen:
            nxt = []
            for u in gen:
                order.append(u)
                for v in succs[u]:
                    indeg[v] -= 1
                    if indeg[v] == 0:
                        nxt.append(v)
            gen = nxt
        return order

    order = topo_order()
    boxes: List[List[int]] = []
    for _ in range(max_path + 1):
        dist = {}
        for v in order:
            best = None
            for u in preds[v]:
                cand = dist[u][0] + weight[(u, v)]
                if best is None or cand > best[0]:
                    best = (cand, u)
            dist[v] = best if (best is not None and best[0] >= 0) else (0, v)
        end = max(dist, key=lambda x: dist[x][0])
        path, u, v = [], None, end
        while u != v:
            path.append(v)
            u = v
            v = dist[v][1]
        path.reverse()
        for a, b in zip(path[:-1], path[1:]):
            weight[(a, b)] = 0.0
        path = [p for p in path if p != 0 and p != last]
        if not path:
            break
        pq, pr = [pair[p][0] for p in path], [pair[p][1] for p in path]
        score = 0.0
        for p in path:
            score += float(val[(p - 1) // top, (p - 1) % top])
        if score > 0:
            qmin, qmax, rmin, rmax = min(pq), max(pq), min(pr), max(pr)
        else:
            qmin = qmax = rmin = rmax = 0
        ave = (rmax - rmin + qmax - qmin) / 2
        ok_len = min(rmax - rmin, qmax - qmin) > min_length
        if ave > 0 and score / ave > min_sim and ok_len and _iou_max([qmin, rmin, qmax, rmax], boxes) < max_iou:
            boxes.append([int(qmin), int(rmin), int(qmax), int(rmax)])
    return boxes


def tn(sims: np.ndarray, tn_max_step=10, tn_top_k=5, max_path=10, min_sim=0.2, min_length=5, max_iou=0.3):
    idx, val = topk_rows(np.asarray(sims), tn_top_k)
    return tn_from_topk(idx, val, tn_max_step, max_path, min_sim, min_length, max_iou)
