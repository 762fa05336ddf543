"""The cp.async row rings of csv_rows_ring / pm_rows_ring (chan_vese_b200/csrc/{csv,pm}_kernels.cu) as a schedule:
which row is requested into which slot when, which commit groups a wait lets through, which slot is read when.

This is a MODEL of the two loops (the kernels themselves are checked against the oracle on the GPU, for many segment
lengths); it exists so that an edit of the ring size, the unroll factor or the wait counts has to keep the invariants
that make the pipeline correct for EVERY segment length n:
  * a slot is read only after the row it is supposed to hold has been requested into it and that request's commit
    group is complete according to the cp.async.wait_group accounting;
  * a slot is not overwritten while the row in it still has a read to come;
  * no row beyond the tail padding of the buffers (TAIL_ROWS) is ever requested.
The ring sizes and the padding are read from the sources."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "chan_vese_b200", "csrc")


def const(fname, name):
    m = re.search(r"constexpr int %s = (\d+);" % name, open(os.path.join(CSRC, fname)).read())
    assert m, (fname, name)
    return int(m.group(1))


class Ring:
    """Slots, commit groups and the checks.  Rows are numbered relative to the first row of the ring."""

    def __init__(self, nslots):
        self.n = nslots
        self.slot_row = [None] * nslots   # row requested into the slot
        self.slot_group = [None] * nslots
        self.open = []                    # slots written since the last commit
        self.committed = 0                # groups committed so far
        self.complete = 0                 # groups known complete (after the last wait)
        self.max_row = -1
        self.last_read = {}               # row -> index of the iteration of its last read (filled by the caller)

    def issue(self, row, slot, reads_left):
        old = self.slot_row[slot]
        assert old is None or reads_left(old) == 0, "slot %d overwritten while row %s still has reads" % (slot, old)
        self.slot_row[slot] = row
        self.slot_group[slot] = None
        self.open.append(slot)
        self.max_row = max(self.max_row, row)

    def commit(self):
        for s in self.open:
            self.slot_group[s] = self.committed
        self.open = []
        self.committed += 1

    def wait(self, pending_allowed):
        self.complete = max(self.complete, self.committed - pending_allowed)

    def read(self, row, slot):
        assert self.slot_row[slot] == row, "slot %d holds row %s, wanted %d" % (slot, self.slot_row[slot], row)
        g = self.slot_group[slot]
        assert g is not None and g < self.complete, "row %d read before its copy is known complete" % row


def csv_schedule(n, ns):
    """csv_rows_ring: slot of row k is k % ns; prologue rows 0..ns-2; iteration r requests row r+ns-1 into the slot of
    row r-1, waits for all but ns-2 groups, reads u of row r+1 and the image of row r."""
    ring = Ring(ns)
    reads = {}  # row -> reads still to come: u part (as S of row k-1, k >= 1) and image part (row k < n); row 0 also as C

    def plan(k):
        c = 0
        if 1 <= k <= n:
            c += 1          # S of iteration k-1
        if k < n:
            c += 1          # image bytes of iteration k
        if k == 0:
            c += 1          # the prologue reads row 0 (C) and may read row 1 (image top)
        return c

    for k in range(0, n + 3 * ns):
        reads[k] = plan(k)
    left = lambda k: reads.get(k, 0)
    for k in range(ns - 1):
        ring.issue(k, k, left)
        ring.commit()
    ring.wait(ns - 3)
    ring.read(0, 0)
    reads[0] -= 1
    ring.read(1, 1 % ns)  # the ra == 0 fix-up reads row 1 early (not counted: it is read again as S)
    for r in range(n):
        ring.issue(r + ns - 1, (r + ns - 1) % ns, left)
        ring.commit()
        ring.wait(ns - 2)
        ring.read(r + 1, (r + 1) % ns)
        reads[r + 1] -= 1
        ring.read(r, r % ns)
        reads[r] -= 1
    assert all(v == 0 for k, v in reads.items() if k <= n), "a planned read did not happen"
    return ring.max_row


def pm_schedule(n, ns):
    """pm_rows_ring: ring row k = image row ra-2+k in slot k % ns; prologue requests rows 0..ns-1 (two per group) and
    consumes rows 0..3; iteration r consumes ring row r+4; every six iterations three pairs request rows r+ns, r+ns+1
    into the slots of rows r, r+1; the tail (< 6 rows) requests nothing."""
    ring = Ring(ns)
    reads = {k: (1 if k <= n + 3 else 0) for k in range(0, n + 4 * ns)}
    left = lambda k: reads.get(k, 0)
    for k in range(0, ns, 2):
        ring.issue(k, k, left)
        ring.issue(k + 1, k + 1, left)
        ring.commit()
    ring.wait(ns // 2 - 2)
    for k in range(4):
        ring.read(k, k)
        reads[k] -= 1
    r = 0
    while r + 6 <= n:
        for j in (0, 2, 4):
            ring.issue(r + j + ns, (r + j) % ns, left)
            ring.issue(r + j + 1 + ns, (r + j + 1) % ns, left)
            ring.commit()
            ring.wait(ns // 2 - 2)
            for q in (r + j + 4, r + j + 5):
                ring.read(q, q % ns)
                reads[q] -= 1
        r += 6
    ring.wait(0)
    while r < n:
        ring.read(r + 4, (r + 4) % ns)
        reads[r + 4] -= 1
        r += 1
    assert all(v == 0 for k, v in reads.items() if k <= n + 3), "a planned read did not happen"
    return ring.max_row


def test_csv_ring_schedule():
    ns, tail = const("csv_kernels.cu", "RING_NS"), const("common.cuh", "TAIL_ROWS")
    for n in range(1, 260):
        top = csv_schedule(n, ns)
        # rows are relative to ra; the buffers end HALO rows after the slab plus TAIL_ROWS (rows n, n+1 are halo rows)
        assert top <= n + 1 + tail, (n, top)


def test_pm_ring_schedule():
    ns, tail, halo = const("pm_kernels.cu", "PM_RING_NS"), const("common.cuh", "TAIL_ROWS"), const("common.cuh", "HALO")
    for n in range(1, 260):
        top = pm_schedule(n, ns)
        # ring row k = image row ra - 2 + k; the last row that exists is rb - 1 + HALO + TAIL_ROWS = ring row n + 1 + halo + tail
        assert top <= n + 1 + halo + tail, (n, top)


def test_the_model_catches_a_broken_schedule():
    """Sanity of the model itself: one group fewer allowed in flight than the kernel waits for must trip it."""
    import pytest

    ns = const("csv_kernels.cu", "RING_NS")
    ring = Ring(ns)
    ring.issue(0, 0, lambda k: 0)
    ring.commit()
    ring.issue(1, 1, lambda k: 0)
    ring.commit()
    ring.wait(1)
    ring.read(0, 0)
    with pytest.raises(AssertionError):
        ring.read(1, 1)           # its group may still be in flight
    with pytest.raises(AssertionError):
        ring.issue(9, 0, lambda k: 1)  # row 0 still has a read to come
