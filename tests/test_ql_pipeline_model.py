"""Model check of the sweep pipeline of k_ql_replay_pipe (svdlibrs_b200/csrc/tridiag_device.cu) on the CPU.

The GPU test (test_ql_replay_on_the_gpu_is_bit_identical) shows that the kernel's result equals the sequential replay, but a
protocol error could hide behind lucky timing.  Here the PROTOCOL itself -- one (sweep id, progress) word per warp, "infinite"
progress until the first batch is done, in-order retirement, "wait until the predecessor has retired or its progress <= the
lowest rotation index of my batch" -- is restated in Python and driven by an adversarial random scheduler over random sweep
lists.  At every batch start the model asserts the safety property the kernel relies on: no EARLIER sweep that is still running
will read or write any column this batch touches (a running sweep owns the columns from its lowest one up to the column it
carries in a register), and that the final z equals the sequential replay (the rotations do not commute).  It also asserts
liveness: some warp can always make progress until every sweep has retired.

Reference for what is replayed: the eigenvector update of imtql2, src/lib.rs:1646-1652.
"""
import numpy as np
import pytest

INF = 0x7FFFFFFF
BATCH = 32


def sequential(n, sweeps):
    z = np.eye(n)
    for hi, rots in sweeps:
        i = hi
        for s, c in rots:
            f = z[:, i + 1].copy()
            z[:, i + 1] = s * z[:, i] + c * f
            z[:, i] = c * z[:, i] - s * f
            i -= 1
    return z


class Warp:
    """One warp of the kernel: replays sweeps w, w + W, ... ; `step()` performs ONE protocol action (or reports that it waits)."""

    def __init__(self, w, W, sweeps, z, slots, owners):
        self.w, self.W, self.sweeps, self.z, self.slots, self.owners = w, W, sweeps, z, slots, owners
        self.v = w          # current sweep
        self.t0 = 0         # rotations of the sweep already replayed
        self.i = None       # next rotation index
        self.pred_retired = False
        self.carry = None   # column held "in a register"
        self.state = "start" if w < len(sweeps) else "done"

    def pred_slot(self):
        return self.slots[(self.w + self.W - 1) % self.W]

    def step(self):
        if self.state == "done":
            return False
        hi, rots = self.sweeps[self.v]
        count = len(rots)
        if self.state == "start":
            self.t0, self.i, self.pred_retired, self.carry = 0, hi, self.v == 0, None
            self.state = "batch" if count else "retire"
            return True
        if self.state == "batch":
            nb = min(BATCH, count - self.t0)
            i_low = self.i - (nb - 1)
            if not self.pred_retired:
                sid, prog = self.pred_slot()
                if sid > self.v - 1:
                    self.pred_retired = True
                elif prog > i_low:
                    return False  # keeps spinning on the predecessor's word
            # ---- safety: no earlier running sweep still owns a column of [i_low, i + 1]
            for u, (lo_u, top_u) in self.owners.items():
                if u < self.v:
                    assert top_u < i_low or lo_u > self.i + 1, (
                        f"sweep {self.v} batch [{i_low}, {self.i + 1}] overlaps running sweep {u} owning [{lo_u}, {top_u}]")
            z = self.z
            if self.t0 == 0:
                self.carry = z[:, self.i + 1].copy()
            for t in range(nb):
                s, c = rots[self.t0 + t]
                x = z[:, self.i].copy()
                z[:, self.i + 1] = s * x + c * self.carry
                self.carry = c * x - s * self.carry
                self.i -= 1
            self.t0 += nb
            lo = hi - count + 1
            self.owners[self.v] = (lo, self.i + 1)  # still to touch: its low end up to the carried column
            if self.t0 < count:
                self.slots[self.w] = (self.v, self.i + 2)  # columns >= i + 2 are final
            else:
                self.state = "final"
            return True
        if self.state == "final":
            self.z[:, self.i + 1] = self.carry
            self.owners.pop(self.v, None)
            self.state = "retire"
            return True
        if self.state == "retire":
            if not self.pred_retired:
                sid, _ = self.pred_slot()
                if sid <= self.v - 1:
                    return False
                self.pred_retired = True
            self.slots[self.w] = (self.v + self.W, INF)
            self.v += self.W
            self.state = "start" if self.v < len(self.sweeps) else "done"
            return True
        raise AssertionError(self.state)


def run_pipeline(n, sweeps, W, rng):
    z = np.eye(n)
    slots = [(w, INF) for w in range(W)]
    owners = {}
    warps = [Warp(w, W, sweeps, z, slots, owners) for w in range(W)]
    # a sweep owns its whole range from the moment it exists until it reports otherwise
    for v, (hi, rots) in enumerate(sweeps):
        if rots:
            owners[v] = (hi - len(rots) + 1, hi + 1)
    steps = 0
    while any(wp.state != "done" for wp in warps):
        order = rng.permutation(W)
        progressed = False
        # adversarial scheduler: usually ONE randomly chosen warp moves; sometimes a burst of one warp
        for w in order:
            burst = int(rng.integers(1, 4)) if rng.random() < 0.3 else 1
            moved = False
            for _ in range(burst):
                if warps[w].step():
                    moved = True
                else:
                    break
            if moved:
                progressed = True
                break
        assert progressed, "deadlock: every warp is waiting"
        steps += 1
        assert steps < 10_000_000
    return z


def random_sweeps(n, nsweeps, rng, kind):
    sweeps = []
    for _ in range(nsweeps):
        if kind == "ql":            # like the QL iteration: ranges that end at a slowly rising lower index
            lo = int(rng.integers(0, max(1, n // 3)))
            hi = int(rng.integers(lo, n - 1))
        elif kind == "short":       # many sweeps shorter than a batch, anywhere
            lo = int(rng.integers(0, n - 1))
            hi = min(n - 2, lo + int(rng.integers(0, 12)))
        else:                       # anything: disjoint, nested, identical ranges
            lo = int(rng.integers(0, n - 1))
            hi = int(rng.integers(lo, n - 1))
        hi = min(hi, n - 2)
        lo = min(lo, hi)
        count = hi - lo + 1 if rng.random() > 0.05 else 0
        th = rng.uniform(0, 2 * np.pi, size=count)
        sweeps.append((hi, list(zip(np.sin(th), np.cos(th)))))
    return sweeps


@pytest.mark.parametrize("kind", ["ql", "short", "any"])
@pytest.mark.parametrize("W", [1, 2, 3, 8])
def test_sweep_pipeline_protocol_is_safe_and_live(kind, W):
    rng = np.random.default_rng([sum(map(ord, kind)), W])
    for trial in range(40):
        n = int(rng.integers(3, 140))
        sweeps = random_sweeps(n, int(rng.integers(1, 40)), rng, kind)
        want = sequential(n, sweeps)
        got = run_pipeline(n, sweeps, W, rng)
        assert np.array_equal(got, want), (kind, W, trial)


def test_the_model_detects_a_broken_protocol(monkeypatch):
    """the check has teeth: with the 'infinite progress until the first batch is done' rule removed (a fresh sweep publishes
    its own top instead), a follower can overtake a predecessor's predecessor, and the model notices"""
    rng = np.random.default_rng(7)
    caught = 0
    orig = Warp.step

    def broken(self):
        was_start = self.state == "start"
        r = orig(self)
        if was_start and self.state == "batch":
            hi, _ = self.sweeps[self.v]
            self.slots[self.w] = (self.v, hi + 2)  # WRONG: claims everything above its own range is final
        return r

    monkeypatch.setattr(Warp, "step", broken)
    for trial in range(200):
        n = int(rng.integers(40, 140))
        sweeps = random_sweeps(n, int(rng.integers(6, 30)), rng, "any")
        try:
            got = run_pipeline(n, sweeps, 4, rng)
            if not np.array_equal(got, sequential(n, sweeps)):
                caught += 1
        except AssertionError:
            caught += 1
    assert caught > 0
