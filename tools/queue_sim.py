"""Timed replay of a work queue of the persistent Cholesky kernel (agp_fused.cu) on the host.

The kernel pops items in queue order, one per free CTA, and spins on dependency counters.  With
per-phase durations calibrated from a traced run (tools/trace_report.py), this replays a queue
without a GPU and reports makespan / waiting time, so queue orders can be compared offline before
they are measured on the device.  Diagnostics only — not part of the product path.
"""
import bisect
import heapq

import numpy as np

DIAG, POTF2, PANEL, PARTIAL, YINIT = 0, 1, 2, 1 << 9, 1 << 10


class Costs:
    """Per-phase durations in microseconds (two CTAs per SM sharing the FP64 pipe)."""

    def __init__(self, prod=18.0, item0=6.0, solve=23.0, potf2=124.0, store=3.0, pop=1.0):
        self.prod = prod      # one 64x128x128 product of the contraction
        self.item0 = item0    # accumulator load + pipeline fill
        self.solve = solve    # triangular solve + store + forward-solve fold
        self.potf2 = potf2
        self.store = store    # store-only epilogue (DIAG, PARTIAL)
        self.pop = pop


def simulate(buf, P, nts, slots=296, costs=None, return_times=False, noise=0.0, seed=0):
    """buf: [n_items][8] int32 as exported by agp_queue_build*.  nts: counter stride (nt_total)."""
    c = costs or Costs()
    rng = np.random.default_rng(seed)
    jit = np.exp(noise * rng.standard_normal(len(buf))) if noise > 0 else np.ones(len(buf))
    n_cnt = 32 + 3 * P * nts + P
    inc = [[] for _ in range(n_cnt)]   # finish times of the increments of every counter (sorted)
    rowdone = lambda p, i: 32 + p * nts + i
    diagu = lambda p, k: 32 + P * nts + p * nts + k
    ppre = lambda p, i: 32 + 2 * P * nts + p * nts + i
    fdone = lambda p: 32 + 3 * P * nts + p

    def reached(cnt, v):
        if v <= 0:
            return 0.0
        lst = inc[cnt]
        assert len(lst) >= v, "queue order violates the producers-first rule"
        return lst[v - 1]

    free = [(0.0, s) for s in range(slots)]
    heapq.heapify(free)
    wait_tot = 0.0
    busy_tot = 0.0
    t_start = np.zeros(len(buf))
    t_end = np.zeros(len(buf))
    for idx, (x, p, k, i, f4, f5, flag, need) in enumerate(buf.tolist()):
        t, partial = x & 0xFF, bool(x & PARTIAL)
        j0, j1, need_k, need_i = f4 & 0xFFFF, f4 >> 16, f5 & 0xFFFF, f5 >> 16
        t0, s = heapq.heappop(free)
        now = t0 + c.pop
        if t == POTF2:
            ready = reached(diagu(p, k), need)
            w = max(0.0, ready - now)
            fin = now + w + c.potf2 * jit[idx]
            bisect.insort(inc[fdone(p)], fin)
        else:
            ready = max(reached(rowdone(p, k), need_k), reached(rowdone(p, i), need_i) if t == PANEL else 0.0)
            if flag >= 0:
                ready = max(ready, reached(flag, need))
            w = max(0.0, ready - now)
            if k == 0 and t == DIAG:
                fin = now + w + 1.0
            else:
                fin = now + w + (c.item0 + c.prod * (j1 - j0)) * jit[idx]
            if t == DIAG:
                fin += c.store
                bisect.insort(inc[diagu(p, k)], fin)
            elif partial:
                fin += c.store
                bisect.insort(inc[ppre(p, i)], fin)
            else:
                rf = reached(fdone(p), k + 1)
                w2 = max(0.0, rf - fin)
                w += w2
                fin += w2 + c.solve * jit[idx]
                bisect.insort(inc[rowdone(p, i)], fin)
        wait_tot += w
        busy_tot += fin - t0 - w
        t_start[idx], t_end[idx] = t0, fin
        heapq.heappush(free, (fin, s))
    makespan = max(t for t, _ in free)
    out = {"makespan_us": makespan, "wait_cta_us": wait_tot, "busy_cta_us": busy_tot,
           "tail_cta_us": sum(makespan - t for t, _ in free), "occupancy": busy_tot / (slots * makespan)}
    if return_times:
        out["t_start"], out["t_end"] = t_start, t_end
    return out


def simulate_dynamic(buf, P, nts, slots=296, costs=None, noise=0.0, seed=0, defer=True, ring_max=1 << 30, return_times=False):
    """Discrete-event replay with the DEFERRAL rule: a CTA that pops an item whose entry
    dependencies are not met parks it in a global ring and pops the next one; every CTA looking for
    work first takes the oldest parked item that has become ready.  defer=False is the plain in-order
    kernel (spin on the popped item)."""
    c = costs or Costs()
    rng = np.random.default_rng(seed)
    n = len(buf)
    jit = np.exp(noise * rng.standard_normal(n)) if noise > 0 else np.ones(n)
    items = buf.tolist()
    cnt = np.zeros(32 + 3 * P * nts + P, dtype=np.int64)
    rowdone = lambda p, i: 32 + p * nts + i
    diagu = lambda p, k: 32 + P * nts + p * nts + k
    ppre = lambda p, i: 32 + 2 * P * nts + p * nts + i
    fdone = lambda p: 32 + 3 * P * nts + p

    def entry_ready(idx):
        x, p, k, i, f4, f5, flag, need = items[idx]
        t = x & 0xFF
        if t == POTF2:
            return cnt[diagu(p, k)] >= need
        j1, need_k, need_i = f4 >> 16, f5 & 0xFFFF, f5 >> 16
        if cnt[rowdone(p, k)] < need_k:
            return False
        if t == PANEL and cnt[rowdone(p, i)] < need_i:
            return False
        return flag < 0 or cnt[flag] >= need

    def durations(idx):
        x, p, k, i, f4, f5, flag, need = items[idx]
        t, partial = x & 0xFF, bool(x & PARTIAL)
        j0, j1 = f4 & 0xFFFF, f4 >> 16
        if t == POTF2:
            return c.potf2 * jit[idx], None
        if t == DIAG:
            return (1.0 if k == 0 else (c.item0 + c.prod * (j1 - j0)) * jit[idx] + c.store), None
        d1 = (c.item0 + c.prod * (j1 - j0)) * jit[idx]
        if partial:
            return d1 + c.store, None
        return d1, c.solve * jit[idx]

    head = 0
    ring = []                       # parked item indices (oldest first)
    ev = [(0.0, 0, s, -1, 0) for s in range(slots)]   # (time, seq, slot, item idx, stage)
    heapq.heapify(ev)
    seq = slots
    idle = []                       # slots with nothing runnable: (slot, since)
    blocked_mid = []                # (slot, idx, since): waiting for potf2 before the solve
    spin = {}                       # in-order mode: slot -> (idx, since)
    wait_tot = busy_tot = 0.0
    last_free = np.zeros(slots)
    t_start = np.zeros(n)
    t_end = np.zeros(n)

    def start(slot, idx, t):
        nonlocal seq, busy_tot
        d1, _ = durations(idx)
        t_start[idx] = t
        heapq.heappush(ev, (t + d1, seq, slot, idx, 1))
        seq += 1
        busy_tot += d1

    def dispatch(slot, t):
        """returns True if the slot got work (or is spinning in in-order mode)"""
        nonlocal head
        if defer:
            for a, idx in enumerate(ring):
                if entry_ready(idx):
                    ring.pop(a)
                    start(slot, idx, t + c.pop)
                    return True
            while head < n:
                idx = head
                if entry_ready(idx):
                    head += 1
                    start(slot, idx, t + c.pop)
                    return True
                if len(ring) >= ring_max:
                    return False
                head += 1
                ring.append(idx)
                t += c.pop * 0.5
            return False
        if head < n:
            idx = head
            head += 1
            if entry_ready(idx):
                start(slot, idx, t + c.pop)
            else:
                spin[slot] = (idx, t)
            return True
        return False

    makespan = 0.0
    while ev:
        t, _, slot, idx, stage = heapq.heappop(ev)
        if idx >= 0:
            x, p, k, i, f4, f5, flag, need = items[idx]
            ty, partial = x & 0xFF, bool(x & PARTIAL)
            if stage == 1 and ty == PANEL and not partial:
                if cnt[fdone(p)] >= k + 1:
                    d2 = durations(idx)[1]
                    busy_tot += d2
                    heapq.heappush(ev, (t + d2, seq, slot, idx, 2))
                    seq += 1
                else:
                    blocked_mid.append((slot, idx, t))
                continue
            # item finished: bump its counter
            if ty == POTF2:
                cnt[fdone(p)] += 1
            elif ty == DIAG:
                cnt[diagu(p, k)] += 1
            elif partial:
                cnt[ppre(p, i)] += 1
            else:
                cnt[rowdone(p, i)] += 1
            t_end[idx] = t
            makespan = max(makespan, t)
            # wake everything that may have become runnable
            if ty == POTF2:
                still = []
                for s2, i2, since in blocked_mid:
                    if cnt[fdone(items[i2][1])] >= items[i2][2] + 1:
                        wait_tot += t - since
                        d2 = durations(i2)[1]
                        busy_tot += d2
                        heapq.heappush(ev, (t + d2, seq, s2, i2, 2))
                        seq += 1
                    else:
                        still.append((s2, i2, since))
                blocked_mid[:] = still
            for s2 in list(spin):
                i2, since = spin[s2]
                if entry_ready(i2):
                    del spin[s2]
                    wait_tot += t - since
                    start(s2, i2, t)
        if not dispatch(slot, t):
            idle.append((slot, t))
        # idle slots retry after every completion
        if idx >= 0 and idle:
            still = []
            for s2, since in idle:
                if s2 == slot and since == t:
                    still.append((s2, since))
                    continue
                if dispatch(s2, t):
                    wait_tot += t - since
                else:
                    still.append((s2, since))
            idle[:] = still
    tail = sum(makespan - since for _, since in idle)
    out = {"makespan_us": makespan, "wait_cta_us": wait_tot, "busy_cta_us": busy_tot, "tail_cta_us": tail}
    if return_times:
        out["t_start"], out["t_end"] = t_start, t_end
    return out
