"""Prototype queue orders for the persistent Cholesky kernel (host-side exploration; the chosen
order is then ported to build_queue in agp_api.cu).  Items use the wire layout of agp_queue_build."""
import numpy as np

DIAG, POTF2, PANEL, PARTIAL, YINIT = 0, 1, 2, 1 << 9, 1 << 10


class Builder:
    def __init__(self, P, nt):
        self.P, self.nt = P, nt
        self.items = []
        self.cov = np.zeros((P, nt, nt), dtype=np.int64)      # contraction coverage of tile (i,k)
        self.n_ppre = np.zeros((P, nt), dtype=np.int64)
        self.n_diag = np.zeros((P, nt), dtype=np.int64)

    def flag_diagu(self, p, k):
        return 32 + self.P * self.nt + p * self.nt + k

    def flag_ppre(self, p, i):
        return 32 + 2 * self.P * self.nt + p * self.nt + i

    def potf2(self, p, k):
        self.items.append((POTF2, p, k, k, 0, 0, -1, int(self.n_diag[p, k])))

    def tile(self, p, i, k, j1):
        """advance tile (i,k) of particle p to coverage j1 (final when j1 == k): both halves"""
        j0 = int(self.cov[p, i, k])
        assert j0 <= j1 <= k and (j1 > j0 or j1 == k)
        final = j1 == k
        for h in (0, 1):
            if i == k:
                flag, need = (self.flag_diagu(p, k), int(self.n_diag[p, k])) if j0 > 0 else (-1, 0)
                self.items.append((DIAG | (h << 8) | (0 if final else PARTIAL), p, k, i, j0 | (j1 << 16), 2 * j1, flag, need))
            else:
                flag, need = (self.flag_ppre(p, i), int(self.n_ppre[p, i])) if j0 > 0 else (-1, 0)
                fl = (0 if final else PARTIAL) | (YINIT if k == 0 else 0)
                self.items.append((PANEL | (h << 8) | fl, p, k, i, j0 | (j1 << 16), (2 * j1) | ((2 * j1) << 16), flag, need))
        if i == k:
            self.n_diag[p, k] += 2
        elif not final:
            self.n_ppre[p, i] += 2
        self.cov[p, i, k] = j1

    def array(self):
        return np.array(self.items, dtype=np.int32)


def order3(P, nt, late=5, split_from=3, bulk_i_major=True, la_pos=(1, 3), diag_pos=(1, 3), potf2_pos=(2, 3), late_lag=1):
    """Early block columns: order 2 (look-ahead split of tiles (k+1,k+1), (k+2,k+1)).  The last `late`
    block columns run right-looking: the trailing triangle is brought up to date when the switch
    column is reached and every later block column only adds single products."""
    b = Builder(P, nt)
    ks = max(nt - late, 1)              # first block column of the late phase
    if nt == 1:
        for p in range(P):
            b.tile(p, 0, 0, 0)
        for p in range(P):
            b.potf2(p, 0)
        return b.array()
    for p in range(P):
        b.tile(p, 0, 0, 0)
    for p in range(P):
        b.potf2(p, 0)
    split = lambda k: split_from <= k < nt
    for k in range(nt - 1):
        latek = k >= ks
        # panels of tile row k+1 first (they feed the next diagonal tile)
        for p in range(P):
            b.tile(p, k + 1, k, k)
        if not latek:
            la = []      # look-ahead partials of block column k+1 over [0,k)
            if split(k + 1) and k >= 1:
                for p in range(P):
                    la.append((p, k + 1, k + 1, k))
                    if k + 2 < nt:
                        la.append((p, k + 2, k + 1, k))
            if k + 1 == ks and k >= 1:
                # switch: bring the rest of the trailing triangle up to [0,k) as well
                for p in range(P):
                    for c in range(k + 1, nt):
                        for i in range(c, nt):
                            if (i, c) not in ((k + 1, k + 1), (k + 2, k + 1)) or not split(k + 1):
                                la.append((p, i, c, k))
            bulk = []
            if bulk_i_major:
                for i in range(k + 2, nt):
                    for p in range(P):
                        bulk.append((p, i, k, k))
            else:
                for p in range(P):
                    for i in range(k + 2, nt):
                        bulk.append((p, i, k, k))
            nb = len(bulk)
            cuts = sorted([(la_pos[0] * nb // la_pos[1], 0), (diag_pos[0] * nb // diag_pos[1], 1), (potf2_pos[0] * nb // potf2_pos[1], 2)])
            pos = 0
            for cut, what in cuts:
                for a in bulk[pos:cut]:
                    b.tile(*a)
                pos = cut
                if what == 0:
                    for a in la:
                        b.tile(*a)
                elif what == 1:
                    for p in range(P):
                        b.tile(p, k + 1, k + 1, k + 1)
                else:
                    for p in range(P):
                        b.potf2(p, k + 1)
            for a in bulk[pos:]:
                b.tile(*a)
        else:
            # late phase, right-looking: every tile below/right is one product behind
            for p in range(P):
                b.tile(p, k + 1, k + 1, k + 1)      # DIAG(k+1) final: [k, k+1)
            for p in range(P):
                b.potf2(p, k + 1)
            for i in range(k + 2, nt):
                for p in range(P):
                    b.tile(p, i, k, k)               # remaining panels of column k: last product + solve
            # trailing updates with block column k:  column k+1 tiles get theirs inside their final items
            # (next phase); tiles further right every `late_lag` columns
            for c in range(k + 2, nt):
                if (c - (k + 1)) % late_lag != 0 and c != k + 2:
                    continue
                for i in range(c, nt):
                    for p in range(P):
                        if b.cov[p, i, c] < k + 1 and c > k + 1:
                            b.tile(p, i, c, k + 1)
    return b.array()
