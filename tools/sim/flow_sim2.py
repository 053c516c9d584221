"""Fluid simulation of one SM running k_whole (d=3, n=30), calibrated on the measured phase timeline:
barrier schedule (as shipped) vs dependency-counter schedule (per-plane / per-a0 counters, whole-function refill).
Development tool; not part of the product."""
import sys, numpy as np

n = 30
inA = lambda a, b, c: a * a + b * b + c * c <= n * n
T = {}
for u in range(n + 1):
    for v in range(n + 1):
        t = sum(1 for k in range(n + 1) if inA(u, v, k))
        if t: T[(u, v)] = t

def terms(t):
    tt = (t + 1) & ~1
    return tt * (tt + 1) // 2

def simulate(chunks, nf=1, exact=True, dataflow=True, nfun=7, lone=0.45, two=0.75, cap=0.93,
             ld=4.2, st=1.2, hst=3.3, fill_cycles=5200.0, fill_lat=800.0, dt=10.0):
    """chunks[h][w]: items of warp w in sweep h.  S0 item (a1,a2), S1 item (a0,a2), S2 item (a0,a1)."""
    nw = len(chunks[0])
    cpt = (4.0 if exact else 2.0) * nf
    needX = np.zeros(n + 1); needY = np.zeros(n + 1)
    for (u, v) in T: needX[v] += 1; needY[u] += 1
    nitems = len(T)
    X = np.zeros((nfun + 1, n + 1)); Y = np.zeros((nfun + 1, n + 1)); Z = np.zeros(nfun + 1)
    fill_done = np.full(nfun + 1, np.inf); fill_done[0] = 0.0
    bar = {}
    st_ = [dict(f=0, h=0, phase='wait', rem=0.0, lat=0.0) for _ in range(nw)]
    def tmax(w, h):
        return max([T[i] for i in chunks[h][w]], default=0)
    t = 0.0
    done = np.zeros(nfun)
    busy = 0.0
    while True:
        for f in range(1, nfun):
            if fill_done[f] == np.inf and Z[f - 1] >= nitems:
                fill_done[f] = t + fill_cycles + fill_lat
        lsu = [w for w in range(nw) if st_[w]['phase'] in ('load', 'store') and st_[w]['lat'] <= 0]
        lrate = 1.0 / max(1, len(lsu))
        comp = {p: sum(1 for w in range(nw) if w % 4 == p and st_[w]['phase'] == 'comp') for p in range(4)}
        fin = True
        arrivals = {}
        for w in range(nw):
            s = st_[w]
            if s['f'] >= nfun: continue
            fin = False
            f, h = s['f'], s['h']
            items = chunks[h][w]
            if s['phase'] == 'wait':
                if dataflow:
                    if h == 0: ok = fill_done[f] <= t
                    elif h == 1: ok = all(X[f, v] >= needX[v] for (u, v) in items)
                    else: ok = all(Y[f, u] >= needY[u] for (u, v) in items)
                else:
                    if h == 0: ok = fill_done[f] <= t
                    else: ok = bar.get((f, h - 1), 0) >= nw
                if ok:
                    s['phase'] = 'load'; s['rem'] = ld * tmax(w, h) * nf; s['lat'] = 60.0
            elif s['phase'] == 'load':
                if s['lat'] > 0: s['lat'] -= dt
                else:
                    s['rem'] -= lrate * dt
                    if s['rem'] <= 0:
                        if h == 2: Z[f] += len(items)
                        s['phase'] = 'comp'; s['rem'] = terms(tmax(w, h)) * cpt
            elif s['phase'] == 'comp':
                k = comp[w % 4]
                rate = min(lone, cap / k) if k != 2 else min(lone, two / 2)
                s['rem'] -= rate * dt; busy += rate * dt
                if s['rem'] <= 0:
                    s['phase'] = 'store'; s['lat'] = 30.0
                    s['rem'] = (st if h < 2 else hst) * tmax(w, h) * nf
            elif s['phase'] == 'store':
                if s['lat'] > 0: s['lat'] -= dt
                else:
                    s['rem'] -= lrate * dt
                    if s['rem'] <= 0:
                        for (u, v) in items:
                            if h == 0: X[f, v] += 1
                            if h == 1: Y[f, u] += 1
                        bar[(f, h)] = bar.get((f, h), 0) + 1
                        s['h'] += 1; s['phase'] = 'wait'
                        if s['h'] == 3:
                            s['h'] = 0; s['f'] += 1; done[f] = max(done[f], t)
        if fin: break
        t += dt
        if t > 3e6: print('timeout'); break
    per = np.diff(done)
    return per, busy / (4 * t)

def make_chunks(nw, nf, ranks):
    """ranks[h][w] = rank (0 = longest) of the chunk warp w works on in sweep h"""
    items = sorted(T.keys(), key=lambda i: -T[i])
    ch = 32 * nf
    base = [items[c * ch:(c + 1) * ch] for c in range(nw)]
    return [[base[ranks[h][w]] for w in range(nw)] for h in range(3)]

def snake(nw):
    # chunk c -> warp; return rank per warp
    r = [0] * nw
    for c in range(nw):
        g = c // 4
        w = g * 4 + ((3 - c % 4) if g % 2 else c % 4)
        r[w] = c
    return r

if __name__ == '__main__':
    for nf, exact, nw, lone, two in ((1, True, 24, 0.45, 0.75), (2, False, 12, 0.5, 0.8)):
        print('fibres/thread', nf, 'exact' if exact else 'fma', 'warps', nw)
        sn = snake(nw)
        same = [sn, sn, sn]
        per, b = simulate(make_chunks(nw, nf, same), nf, exact, dataflow=False, lone=lone, two=two)
        print('  barriers, same rank in every sweep     : period', per[-2:], 'pipe', round(b, 3))
        per, b = simulate(make_chunks(nw, nf, same), nf, exact, dataflow=True, lone=lone, two=two)
        print('  counters, same rank in every sweep     : period', per[-2:], 'pipe', round(b, 3))
        for sh in ((0, nw // 3, 2 * nw // 3), (0, nw // 2, 0), (0, 2 * nw // 3, nw // 3)):
            # rotation within the same sub-partition: shift ranks by multiples of 4 keeps w % 4 classes balanced only roughly
            ranks = [[(sn[w] + sh[h]) % nw for w in range(nw)] for h in range(3)]
            per, b = simulate(make_chunks(nw, nf, ranks), nf, exact, dataflow=True, lone=lone, two=two)
            print('  counters, ranks rotated by', sh, ': period', per[-2:], 'pipe', round(b, 3))
        rev = [sn, [nw - 1 - x for x in sn], sn]
        per, b = simulate(make_chunks(nw, nf, rev), nf, exact, dataflow=True, lone=lone, two=two)
        print('  counters, S1 reversed                  : period', per[-2:], 'pipe', round(b, 3))
