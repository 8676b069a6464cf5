"""Model check of the mirrored-halo hand-offs (jacobi.cuh TbSync, fluid.cu run_jacobi): ranks are threads, a
launch is a few randomly delayed events, the ghost rows of the two ping-pong buffers carry (level) tags.
The protocol under test, per solve s with n >= 2 blocked launches on every rank:
  * launch n-1 ("penultimate") is the last reader of the ghost rows of the buffer launch n writes; when its grid
    is done it stores s into the neighbours' "done reading" word;
  * launch n waits for "done reading" >= s, stores its boundary rows into the neighbours' ghost rows of ITS
    output buffer, then stores s into the neighbours' "mirror ready" word;
  * launch 1 of solve s+1 waits for "mirror ready" >= s before it reads ghost rows.
Invariants asserted with the model's own ground truth (never with the flags): a neighbour's ghost rows are never
overwritten while one of its launches may still read them, and launch 1 always finds the previous solve's rows."""
import random
import threading
import time

import pytest


class Rank:
    def __init__(self):
        self.done_reading = [0, 0]      # flag words, index = side the writer sits on (0 below, 1 above)
        self.mirror_ready = [0, 0]
        self.ghost = [[(0, 0), (0, 0)], [(0, 0), (0, 0)]]   # [buffer][side] -> (solve, level) of the rows stored there
        self.reading = [[0, 0], [0, 0]] # [buffer][side] -> launches currently reading those ghost rows
        self.lock = threading.Lock()


def run_rank(r, ranks, world, solves, nlaunch, rng, errors):
    me = ranks[r]
    nbr = [(r - 1, 0), (r + 1, 1)]      # (rank, side as seen from me)
    nbr = [(q, s) for q, s in nbr if 0 <= q < world]
    spin = lambda cond: [time.sleep(0) for _ in iter(lambda: not cond(), False)]
    cur = 0                              # buffer holding the input of the next launch
    try:
        for s in range(1, solves + 1):
            for k in range(1, nlaunch + 1):
                src, dst = cur, cur ^ 1
                if k == 1 and s > 1:     # wait for the rows the neighbours mirrored at the end of solve s-1
                    for q, side in nbr:
                        spin(lambda side=side: me.mirror_ready[side] >= s - 1)
                # --- the launch reads the ghost rows of src on both sides
                for q, side in nbr:
                    with me.lock:
                        me.reading[src][side] += 1
                        tag = me.ghost[src][side]
                    if k == 1 and s > 1 and tag != (s - 1, nlaunch):
                        errors.append(f"rank {r} solve {s}: first launch read ghost rows {tag}, expected {(s - 1, nlaunch)}")
                time.sleep(rng.random() * 2e-4)
                for q, side in nbr:
                    with me.lock:
                        me.reading[src][side] -= 1
                if k == nlaunch - 1:     # grid done: last reader of the rows launch n of the neighbours will overwrite
                    for q, side in nbr:
                        ranks[q].done_reading[side ^ 1] = s
                if k == nlaunch:         # mirror my boundary rows into the neighbours' ghost rows of dst
                    for q, side in nbr:
                        spin(lambda side=side: me.done_reading[side] >= s)
                        other = ranks[q]
                        with other.lock:
                            if other.reading[dst][side ^ 1]:
                                errors.append(f"rank {r} solve {s}: overwrote ghost rows rank {q} was still reading")
                            other.ghost[dst][side ^ 1] = (s, nlaunch)
                        time.sleep(rng.random() * 1e-4)
                        other.mirror_ready[side ^ 1] = s
                cur = dst
                time.sleep(rng.random() * 1e-4)
    except Exception as e:              # pragma: no cover
        errors.append(repr(e))


@pytest.mark.parametrize("world,nlaunch", [(2, 2), (2, 5), (3, 3), (4, 4)])
def test_mirrored_halo_handoffs_never_race(world, nlaunch):
    for trial in range(3):
        ranks = [Rank() for _ in range(world)]
        errors = []
        ts = [threading.Thread(target=run_rank, args=(r, ranks, world, 25, nlaunch, random.Random(100 * trial + r), errors))
              for r in range(world)]
        for t in ts:
            t.start()
        for t in ts:
            t.join(timeout=60)
        assert not any(t.is_alive() for t in ts), "deadlock"
        assert not errors, errors[:3]
