"""BPP-k lookahead for MANY sequences at once (second half).

The reference decides where to put the current item when the next k-1 items are known by a small tree search over the
orders in which the k items could be placed *virtually* (`acktr/reorder.py:53-262`, driven by `unified_test.py:9-28`):
up to `min(100, (k-1)!)` simulations, each a chain of (network evaluation, env step) pairs on a deep copy of the env.
The simulations of one search depend on each other through the tree's statistics, so they cannot run side by side --
but the searches of different sequences are independent and every one of them alternates the same two requests.  Here
all searches advance in lockstep: one tick = ONE batched network evaluation + ONE batched placement step for every
sequence that is still searching.  The tree logic stays on the host (a generator per sequence, restating the
reference's rules below); the device does what it is good at -- the convolutions and `Space.check_box` /
`update_height_graph` for the whole batch (`BinPackingVecEnv.import_state` -> `step_device` -> `export_state`, the same
engine calls the single-env view makes for one bin).

Reference semantics kept (file:line = acktr/reorder.py):
  * node bookkeeping `Node` :7-50 -- value = best return seen below the node; a child inherits its parent's value until it
    has one of its own; exhausted (`visit >= height!`) and disabled subtrees score -1000 / -10000; exploration bonus
    0.5 * sqrt(parent.visit) / (visit + 1); only children scoring > 0 are ever selected (:135-143);
  * the item arriving last may only be placed last (:129-131);
  * virtual placement out of arrival order: later-arriving items that were already placed are hidden from the network by
    raising their footprint cells to the bin height (`get_mixed_obs` :92-98) and remembered in per-item masks
    (`update_mask` :100-104); a fully raised map ends the simulation (`will_terminate` :106-112);
  * a simulation that ends early is discarded when its order is inconsistent (:171-180), else it reports the return so
    far; a complete one reports return + the critic's value (:183-190); the action reported upward is the position chosen
    for item 0, or the literal 0 when item 0 itself ended the episode (:178, sic);
  * `get_baseline` :226-243 (arrival order, greedy) seeds the root; the search result replaces it only if it is better
    by `v_bound` = 0.1 (:256-260);
  * positions come from `nnModel.evaluate` (`acktr/model_loader.py:39-68`): argmax of softmax(logits) * (predicted mask >= 0.5),
    taken as `np.argsort(...)[-1]` like the reference (:121).
Rotation is not supported by the reference's lookahead (mask length L*W, :62-63) and not here either.
"""
import math

import numpy as np
import torch

P_BOUND = 0.8          # reorder.py:54 -> pos_num = int(1 / p_bound) = 1
V_BOUND = 0.1          # reorder.py:54


class _Node(object):
    __slots__ = ("best", "action", "item", "parent", "height", "children", "open_children", "visits", "capacity")

    def __init__(self, parent, item, height):
        self.best, self.action = None, None
        self.item, self.parent, self.height = item, parent, height
        self.children = []
        self.open_children = height                       # dis_num
        self.visits = 0
        self.capacity = math.factorial(height) if height != -1 else 1

    def exploit(self):
        if self.best is None:
            return self.parent.exploit()
        if self.visits >= self.capacity:
            return -1000
        if self.open_children <= 0:
            return -10000
        return self.best

    def score(self):
        return self.exploit() + 0.5 * np.sqrt(self.parent.visits) / (self.visits + 1)

    def close(self):
        self.open_children = 0
        if self.parent is None:
            return
        self.parent.open_children -= 1
        if self.parent.open_children == 0:
            self.parent.close()

    def report(self, value, action):
        node = self
        while node is not None:
            node.visits += 1
            if node.best is None or value > node.best:
                node.best, node.action = value, action
            node = node.parent


class _Bin(object):
    """Host-side state of one sequence during a search: the real state (root) and the state of the running simulation."""

    def __init__(self, plain, counter, vol, boxes, size):
        self.size = size
        self.root = (np.array(plain, dtype=np.int32).reshape(size[0], size[1]), int(counter), int(vol))
        self.boxes = [tuple(int(v) for v in b) for b in boxes]
        self.rewind()

    def rewind(self):
        self.plain, self.counter, self.vol = self.root[0].copy(), self.root[1], self.root[2]

    def observation(self, box):
        """cur_observation (envs/bpp0/bin3D.py:61-66): int32 [4*L*W]."""
        planes = [self.plain] + [np.full(self.plain.shape, v, dtype=np.int32) for v in box]
        return np.stack(planes).reshape(-1)

    def reward(self, box):
        L, W, H = self.size
        return (box[0] * box[1] * box[2]) / (L * W * H) * 10         # bin3D.py:44-46,121


def _search(b, times):
    """Generator for one sequence.  Yields ('eval', obs int32[4A]) and receives (value, scores float[A]); yields
    ('step', box, pos) and receives (done, plain_after int32[L,W]).  Returns (action, value, default)."""
    k = len(b.boxes)
    L, W, H = b.size
    area = L * W

    # ---- baseline: arrival order, greedy (reorder.py:226-243)
    b.rewind()
    nor_exp, nor_act = 0, None
    for i in range(k):
        val, scores = yield ("eval", b.observation(b.boxes[i]))
        act = int(np.argmax(scores))
        if i == 0:
            nor_act = act
        done, plain = yield ("step", b.boxes[i], act)
        if done:
            break
        b.plain = plain
        if i == k - 1:
            nor_exp += val
            break
        nor_exp += b.reward(b.boxes[i])

    root = _Node(None, None, k - 1)
    root.best, root.action = nor_exp, nor_act
    for _ in range(min(times, math.factorial(k - 1))):
        b.rewind()
        remaining = list(range(k))
        masks = np.ones((k, area))
        node, value, action = root, 0, None
        while True:
            if not node.children:
                for idx in remaining:
                    if idx == k - 1 and len(remaining) > 1:
                        continue
                    node.children.append(_Node(node, idx, node.height - 1))
            nxt, nxt_score = None, 0
            for child in node.children:
                s = child.score()
                if s > nxt_score:
                    nxt, nxt_score = child, s
            if nxt is None:
                node.report(-10000, None)
                break
            idx = nxt.item
            box = b.boxes[idx]
            # what the network sees: cells of later-arriving items already placed are raised to the bin height
            obs = b.observation(box).reshape(4, -1)
            mixed = obs[0].astype(np.float64)
            for i in range(idx + 1, k):
                mixed = H * (1 - masks[i]) + mixed * masks[i]
            obs[0] = mixed
            val, scores = yield ("eval", obs.reshape(-1))
            pos = int(np.argsort(scores)[-1])
            full = np.sum(mixed) == area * H
            done, plain = yield ("step", box, pos)
            if done or full:
                inconsistent = any((i in remaining and i < idx) or (i not in remaining and i >= idx) for i in range(k))
                if inconsistent:
                    nxt.close()
                    node.report(-10000, None)
                elif idx == 0:
                    node.report(value + 0, 0)
                else:
                    node.report(value + 0, action)
                break
            b.plain = plain
            if len(remaining) == 1:
                node.report(value + val, pos if idx == 0 else action)
                break
            remaining.remove(idx)
            cm = masks[idx].reshape(L, W).copy()
            cm[pos // W:pos // W + box[0], pos % W:pos % W + box[1]] = 0
            masks[idx] = cm.reshape(-1)
            value = value + b.reward(box)
            if action is None and idx == 0:
                action = pos
            node = nxt
    best, act = root.best, root.action
    if act != nor_act and best - nor_exp < V_BOUND:
        best, act = nor_exp, nor_act
    return act, best, act == nor_act


def reorder_search_batch(policy, engine, plains, counters, vols, box_lists, times=100):
    """One BPP-k decision for each of n sequences.

    policy     `b200bpp.PackingPolicy` (or anything with `evaluate(obs float32[n,4A]) -> (value[n], scores[n,A])`)
    engine     an n-bin `BinPackingVecEnv` used as the placement engine (any generator; its own items are never used),
               or an object with its `import_state / step_device / export_state / device` surface
    plains     int [n, L*W] current heightmaps; counters / vols int [n] (len(space.boxes), placed volume)
    box_lists  n lists of the k known items, arrival order (`box_creator.preview(k)`)
    -> (actions int64[n], values float64[n], default bool[n]) as `ReorderTree.reorder_search()` returns them per sequence.
    """
    n = len(box_lists)
    size = tuple(int(v) for v in engine.bin_size)
    L, W, _ = size
    A = L * W
    dev = engine.device
    bins = [_Bin(np.asarray(plains[i]), counters[i], vols[i], box_lists[i], size) for i in range(n)]
    gens = [_search(b, times) for b in bins]
    pending = [g.send(None) for g in gens]                 # every search starts with an 'eval' request
    results = [None] * n
    obs_h = np.zeros((n, 4 * A), dtype=np.float32)
    plain_h = np.zeros((n, A), dtype=np.int32)
    box_h = np.ones((n, 3), dtype=np.int32)
    cnt_h, vol_h = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
    act_h = np.zeros(n, dtype=np.int64)
    while True:
        live = [i for i in range(n) if results[i] is None]
        if not live:
            break
        # ---- tick, first half: one batched evaluation
        for i in live:
            assert pending[i][0] == "eval"
            obs_h[i] = pending[i][1]
        with torch.no_grad():
            value, scores = policy.evaluate(torch.from_numpy(obs_h).to(dev))
        value, scores = value.detach().cpu().numpy(), scores.detach().cpu().numpy()
        for i in live:
            pending[i] = gens[i].send((float(value[i]), scores[i]))
        # ---- second half: one batched placement (check_box + update_height_graph) on the simulations' states
        for i in live:
            kind, box, pos = pending[i]
            assert kind == "step"
            plain_h[i], box_h[i], act_h[i] = bins[i].plain.reshape(-1), box, pos
            cnt_h[i], vol_h[i] = bins[i].counter, bins[i].vol
        engine.import_state(plain=plain_h, box=box_h, counter=cnt_h, vol=vol_h)
        _, _, done, _ = engine.step_device(torch.from_numpy(act_h).to(dev))
        done = done.detach().cpu().numpy()
        new_plain = engine.export_state()[0].detach().cpu().numpy().reshape(n, L, W)
        for i in live:
            try:
                pending[i] = gens[i].send((bool(done[i]), new_plain[i].astype(np.int32)))
            except StopIteration as fin:
                results[i] = fin.value
    return (np.array([r[0] for r in results], dtype=np.int64), np.array([r[1] for r in results], dtype=np.float64),
            np.array([bool(r[2]) for r in results]))
