"""Fluid simulation of one SM running the whole-function kernel (d=3): barrier schedule vs dataflow schedule.

Resources: 4 FP64 pipes (one per sub-partition, processor sharing among the warps that are in their compute phase,
a lone warp reaches at most RMAX of the pipe), one LSU (1 wavefront per cycle, shared), fill engine (latency + bandwidth).
Development tool for whole_kernel design decisions; not part of the product.
"""
import sys, math, itertools
import numpy as np

n = 30
def geometry(n):
    inA = lambda a, b, c: a * a + b * b + c * c <= n * n
    t01 = {}  # S2 items (a0,a1) -> length
    for a0 in range(n + 1):
        for a1 in range(n + 1):
            t = sum(1 for k in range(n + 1) if inA(a0, a1, k))
            if t: t01[(a0, a1)] = t
    return t01
T = geometry(n)           # by symmetry the same table serves all three sweeps: item (u,v) -> length
NIT = len(T)

def terms(t, rows=2):
    tt = ((t + rows - 1) // rows) * rows
    return tt * (tt + 1) // 2

class Warp:
    pass

def build(assign, exact=True, dataflow=True, nwarps=23):
    """assign: function (sweep h) -> list of chunks (lists of 32 items (u,v)) in warp order 0..nwarps-1"""
    return assign

def simulate(chunks, smsp_of, exact=True, dataflow=True, nfun=6, lsu_scale=1.0, fill_lat=1500.0, fill_bw=23.0,
             rmax=0.86, ptot=0.93, verbose=False):
    # chunks[h][w] = list of items (u,v): sweep 0: item (a1,a2); sweep 1: (a0,a2); sweep 2: (a0,a1)
    nw = len(chunks[0])
    cyc_per_term = 4.0 if exact else 2.0
    # dependency groups: S0 item (a1,a2): waits fill[a1], signals X[a2]; S1 (a0,a2): waits X[a2], signals Y[a0];
    # S2 (a0,a1): waits Y[a0], signals (after load) Z[a1]
    needX = np.zeros(n + 1); needY = np.zeros(n + 1); needZ = np.zeros(n + 1)
    for (u, v) in T:
        needX[v] += 1   # S0 items by a2
        needY[u] += 1   # S1 items by a0
        needZ[v] += 1   # S2 items by a1
    slab_bytes = np.zeros(n + 1)
    for (a1, a2), t in T.items(): slab_bytes[a1] += 8 * t
    X = np.zeros((nfun + 1, n + 1)); Y = np.zeros((nfun + 1, n + 1)); Z = np.zeros((nfun + 1, n + 1))
    fill_done = np.full((nfun + 1, n + 1), np.inf)
    fill_done[0, :] = 0.0
    fill_issued = np.zeros((nfun + 1, n + 1), bool); fill_issued[0, :] = True
    fill_free_at = 0.0
    # warp state
    st = []
    for w in range(nw):
        st.append(dict(f=0, h=0, phase='wait', rem=0.0, lat=0.0))
    def work(w, h):
        items = chunks[h][w]
        tm = max(T[i] for i in items) if items else 0
        return terms(tm) * cyc_per_term, tm, len(items)
    t = 0.0; dt = 8.0
    done_time = np.zeros(nfun)
    bar_count = {}
    log = []
    pipe_busy = 0.0
    while True:
        # fill engine (dataflow: per slab when Z complete; barrier: whole function when all Z complete)
        for f in range(1, nfun):
            if dataflow:
                for s in range(n + 1):
                    if not fill_issued[f, s] and Z[f - 1, s] >= needZ[s]:
                        start = max(t, fill_free_at)
                        fill_free_at = start + slab_bytes[s] / fill_bw
                        fill_done[f, s] = fill_free_at + fill_lat
                        fill_issued[f, s] = True
            else:
                if not fill_issued[f, 0] and all(Z[f - 1, s] >= needZ[s] for s in range(n + 1)):
                    start = max(t, fill_free_at)
                    fill_free_at = start + slab_bytes.sum() / fill_bw
                    fill_done[f, :] = fill_free_at + fill_lat * 0.3
                    fill_issued[f, :] = True
        # advance phases
        lsu_users = [w for w in range(nw) if st[w]['phase'] in ('load', 'store') and st[w]['lat'] <= 0]
        lsu_rate = min(1.0, 1.0 * 1.0) / max(1, len(lsu_users))
        comp = {p: [w for w in range(nw) if smsp_of[w] == p and st[w]['phase'] == 'comp'] for p in range(4)}
        allfin = True
        for w in range(nw):
            s = st[w]
            if s['f'] >= nfun: continue
            allfin = False
            f, h = s['f'], s['h']
            items = chunks[h][w]
            if s['phase'] == 'wait':
                ok = True
                if dataflow:
                    for (u, v) in items:
                        if h == 0 and not (fill_done[f, u] <= t): ok = False; break
                        if h == 1 and X[f, v] < needX[v]: ok = False; break
                        if h == 2 and Y[f, u] < needY[u]: ok = False; break
                else:
                    key = (f, h)
                    if h == 0: ok = all(fill_done[f, :] <= t)
                    elif h == 1: ok = all(X[f, :] >= needX)
                    else: ok = all(Y[f, :] >= needY)
                if ok or not items:
                    wk, tm, ni = work(w, h)
                    s['phase'] = 'load'; s['rem'] = 2.0 * tm * lsu_scale; s['lat'] = 60.0
                    if not items: s['rem'] = 0
            elif s['phase'] == 'load':
                if s['lat'] > 0: s['lat'] -= dt
                else:
                    s['rem'] -= lsu_rate * dt
                    if s['rem'] <= 0:
                        if h == 2:
                            for (u, v) in items: Z[f, v] += 1
                        s['phase'] = 'comp'; s['rem'] = work(w, h)[0]
            elif s['phase'] == 'comp':
                k = len(comp[smsp_of[w]])
                rate = min(rmax, ptot / k)
                s['rem'] -= rate * dt; pipe_busy += rate * dt
                if s['rem'] <= 0:
                    tm = work(w, h)[1]
                    s['phase'] = 'store'; s['lat'] = 30.0
                    s['rem'] = (2.0 * tm * lsu_scale) if h < 2 else 10.0 * tm  # HBM store: ~10 sectors per instr at 1 sector/clk
            elif s['phase'] == 'store':
                if s['lat'] > 0: s['lat'] -= dt
                else:
                    s['rem'] -= lsu_rate * dt
                    if s['rem'] <= 0:
                        for (u, v) in items:
                            if h == 0: X[f, v] += 1
                            if h == 1: Y[f, u] += 1
                        s['h'] += 1; s['phase'] = 'wait'
                        if s['h'] == 3:
                            s['h'] = 0; s['f'] += 1
                            done_time[f] = max(done_time[f], t)
        if allfin: break
        t += dt
        if t > 2e6: print('timeout'); break
    per = np.diff(done_time)
    return done_time, per, pipe_busy / (4 * t)

def chunks_sorted(rot=(0, 0, 0), nw=23, rev=(False, False, False)):
    items = sorted(T.keys(), key=lambda i: -T[i])
    base = [items[c * 32:(c + 1) * 32] for c in range(nw)]
    out = []
    for h in range(3):
        ch = [None] * nw
        for w in range(nw):
            r = (w + rot[h]) % nw
            if rev[h]: r = nw - 1 - r
            ch[w] = base[r]
        out.append(ch)
    return out

if __name__ == '__main__':
    nw = 23
    snake = [ (w % 4) if (w // 4) % 2 == 0 else 3 - (w % 4) for w in range(nw)]
    for exact in (True, False):
        print('exact' if exact else 'fma')
        ch = chunks_sorted()
        d, per, pb = simulate(ch, snake, exact=exact, dataflow=False)
        print('  barrier, same rank every sweep: period', per[-3:], 'pipe busy', round(pb, 3))
        d, per, pb = simulate(ch, snake, exact=exact, dataflow=True)
        print('  dataflow, same rank every sweep: period', per[-3:], 'pipe busy', round(pb, 3))
        for rot in ((0, 8, 16), (0, 12, 0), (0, 6, 12)):
            ch = chunks_sorted(rot=rot)
            d, per, pb = simulate(ch, snake, exact=exact, dataflow=True)
            print('  dataflow, rot', rot, ': period', per[-3:], 'pipe busy', round(pb, 3))
        ch = chunks_sorted(rev=(False, True, False))
        d, per, pb = simulate(ch, snake, exact=exact, dataflow=True)
        print('  dataflow, rev S1: period', per[-3:], 'pipe busy', round(pb, 3))
