"""Prototype: build the work queue of the persistent Cholesky kernel by list scheduling.

Simulates the CTAs popping items (calibrated durations, tools/queue_sim.py) and, every time a CTA
pair frees up, enqueues the READY tile with the longest remaining critical path (HLFET), splitting
contractions dynamically: a tile whose operands are final only up to block column a < k gets a
store-only PARTIAL item over [cov, a) if that is at least `min_chunk` products.  The result is an
ordinary in-order queue (producers always earlier), replayable by tests/test_abi_host._replay_queue.
Diagnostics / design exploration only.
"""
import heapq

import numpy as np

DIAG, POTF2, PANEL, PARTIAL, YINIT = 0, 1, 2, 1 << 9, 1 << 10
INF = 1e30


def build(P, nt, slots=296, costs=None, min_chunk=3, chain=None):
    from queue_sim import Costs

    c = costs or Costs(prod=15.3, item0=6.5, solve=24.0, potf2=107.0, store=3.3, pop=0.8)
    nts = nt
    ppre_idx = lambda p, i: 32 + 2 * P * nts + p * nts + i
    diagu_idx = lambda p, k: 32 + P * nts + p * nts + k
    # per-column critical chain: potf2 -> last product + solve of panel (k+1,k) -> last product of diag(k+1)
    col_chain = chain or (c.potf2 + (c.item0 + c.prod + c.solve) + (c.item0 + c.prod + c.store))
    row_step = c.item0 + c.prod + c.solve

    ii, kk = np.meshgrid(np.arange(nt), np.arange(nt), indexing="ij")
    lower = kk <= ii
    # bottom level of the FINAL item of tile (i,k) (excluding its own remaining contraction)
    bl_tile = np.where(kk < ii, (nt - ii) * col_chain + (ii - kk - 1) * row_step + c.item0 + c.prod + c.store + c.potf2,
                       (nt - kk) * col_chain)          # diagonal tile: feeds potf2(k)
    cov = np.zeros((P, nt, nt), dtype=np.int64)        # contraction coverage (block columns)
    tcov = np.zeros((P, nt, nt))                       # est. finish of the last item on the tile
    final = np.zeros((P, nt, nt), dtype=bool)          # final item enqueued
    final[:, ~lower] = True
    tfin = np.full((P, nt, nt), INF)                   # est. finish of the final PANEL (i,k) (both halves)
    nrow = np.zeros((P, nt), dtype=np.int64)           # leading block columns of row i with final panels enqueued
    tpot = np.full((P, nt), INF)                       # est. finish of POTF2(k)
    npot = np.zeros(P, dtype=np.int64)                 # potf2 items enqueued
    n_ppre = np.zeros((P, nt), dtype=np.int64)         # PARTIAL panel items enqueued per tile row
    n_diag = np.zeros((P, nt), dtype=np.int64)         # DIAG items enqueued per diagonal tile
    # rowready[p,i,a] = est. time the panels (i, 0..a-1) are all final
    rowready = np.full((P, nt, nt + 1), INF)
    rowready[:, :, 0] = 0.0

    free = [(0.0, s) for s in range(slots)]
    heapq.heapify(free)
    items = []

    def push(t, h, p, k, i, j0, j1, partial, flag, need):
        flags = PARTIAL if partial else 0
        if t == PANEL and k == 0:
            flags |= YINIT
        f5 = 0 if t == POTF2 else (2 * j1) | ((2 * j1 if t == PANEL else 0) << 16)
        items.append((t | (h << 8) | flags, p, k, i, j0 | (j1 << 16), f5, flag, need))

    total_tiles = int(lower.sum()) * P
    done_tiles = 0
    pots_left = P * nt
    while done_tiles < total_tiles or pots_left > 0:
        t0 = free[0][0]
        # ---- potf2 candidates: all DIAG items of (k,k) enqueued -----------------------------
        kq = npot.copy()
        has = kq < nt
        kq_c = np.minimum(kq, nt - 1)
        pidx = np.arange(P)
        pot_ok = has & final[pidx, kq_c, kq_c]
        pot_ready = np.where(pot_ok, tcov[pidx, kq_c, kq_c], INF)
        pot_bl = np.where(pot_ok, (nt - kq_c) * col_chain, -1.0)
        # ---- tile candidates ------------------------------------------------------------------
        # available range end (producers ENQUEUED): min(nrow[i], nrow[k], k)
        a_enq = np.minimum(np.minimum(nrow[:, :, None], nrow[:, None, :]), kk[None])
        # producers FINISHED by t0: largest a with rowready[i][a], rowready[k][a] <= t0
        rr = (rowready <= t0 + 1e-9)
        nready = rr.sum(axis=2) - 1                     # leading columns final by now (rowready monotone in practice)
        a_now = np.minimum(np.minimum(nready[:, :, None], nready[:, None, :]), kk[None])
        a_now = np.minimum(a_now, a_enq)
        own_ok = tcov <= t0 + 1e-9
        # a final item also needs potf2(k) enqueued (panel) — the solve waits inside the item
        can_final = (a_now == kk[None]) & ((npot[:, None, None] > kk[None]) | (ii == kk)[None])
        span = a_now - cov
        cand = (~final) & own_ok & ((can_final) | ((span >= min_chunk) & (a_now < kk[None])))
        rem = (kk[None] - cov) * c.prod
        bl = np.where(cand, bl_tile[None] + rem, -1.0)
        best = int(np.argmax(bl))
        best_bl = bl.flat[best]
        bp = int(np.argmax(np.where(pot_ready <= t0 + 1e-9, pot_bl, -1.0)))
        bp_bl = pot_bl[bp] if pot_ready[bp] <= t0 + 1e-9 else -1.0
        if best_bl < 0 and bp_bl < 0:
            # nothing ready: advance time to the next event (skip this slot forward)
            nxt = []
            if pot_ok.any():
                nxt.append(pot_ready[pot_ok].min())
            fut = rowready[rowready > t0 + 1e-9]
            if fut.size:
                nxt.append(fut.min())
            fc = tcov[(~final) & (tcov > t0 + 1e-9)]
            if fc.size:
                nxt.append(fc.min())
            fp = tpot[(tpot > t0 + 1e-9) & (tpot < INF)]
            if fp.size:
                nxt.append(fp.min())
            tn = min(nxt)
            # idle every slot that is free before tn up to tn
            newfree = [(max(t, tn), s) for t, s in free]
            heapq.heapify(newfree)
            free = newfree
            continue
        if bp_bl >= best_bl:
            p = bp
            k = int(npot[p])
            t_s, s = heapq.heappop(free)
            fin = t_s + c.pop + c.potf2
            push(POTF2, 0, p, k, k, 0, 0, False, -1, int(n_diag[p, k]))
            tpot[p, k] = fin
            npot[p] += 1
            pots_left -= 1
            heapq.heappush(free, (fin, s))
            continue
        p, i, k = np.unravel_index(best, bl.shape)
        p, i, k = int(p), int(i), int(k)
        j0, j1 = int(cov[p, i, k]), int(a_now[p, i, k])
        is_final = j1 == k
        fins = []
        for h in (0, 1):
            t_s, s = heapq.heappop(free)
            now = t_s + c.pop
            if i == k:
                flag, need = (diagu_idx(p, k), int(n_diag[p, k])) if j0 > 0 else (-1, 0)
                fin = now + (1.0 if k == 0 else c.item0 + c.prod * (j1 - j0) + c.store)
                push(DIAG, h, p, k, i, j0, j1, not is_final, flag, need)
            else:
                flag, need = (ppre_idx(p, i), int(n_ppre[p, i])) if j0 > 0 else (-1, 0)
                fin = now + c.item0 + c.prod * (j1 - j0)
                if is_final:
                    fin = max(fin, tpot[p, k]) + c.solve
                else:
                    fin += c.store
                push(PANEL, h, p, k, i, j0, j1, not is_final, flag, need)
            fins.append(fin)
            heapq.heappush(free, (fin, s))
        fin = max(fins)
        # NOTE: counters are bumped after BOTH halves were pushed, so the two halves of one tile carry the same need
        if i == k:
            n_diag[p, k] += 2
        elif not is_final:
            n_ppre[p, i] += 2
        cov[p, i, k] = j1
        tcov[p, i, k] = fin
        if is_final:
            final[p, i, k] = True
            done_tiles += 1
            if i != k:
                tfin[p, i, k] = fin
                nrow[p, i] += 1
                assert nrow[p, i] == k + 1, "panels of a tile row become final in column order"
                rowready[p, i, k + 1] = max(rowready[p, i, k], fin)
            else:
                # row k has k panels: all final already (needed for a_now == k)
                pass
    return np.array(items, dtype=np.int32)
