"""List-scheduling model of the one-warp-CTA marching launch: CTA cost = width + c steps, `slots` resident CTAs,
CTAs dispatched in launch order (strip fastest, chunk-column slowest).  Fit against the measured sweeps, then compare
uniform and tapered chunk-width sequences."""
import heapq, sys, math
import numpy as np
def makespan(widths, strips, slots, c):
    free = [0.0] * slots
    heapq.heapify(free)
    end = 0.0
    for w in widths:
        for _ in range(strips):
            t = heapq.heappop(free)
            t2 = t + w + c
            end = max(end, t2)
            heapq.heappush(free, t2)
    return end
def uniform(ncol, Mx):
    n = math.ceil(ncol / Mx)
    return [min(Mx, ncol - k * Mx) for k in range(n)]
meas = {  # CD, median us (slab3)
 128: {4: 66.2, 5: 63.9, 6: 65.3, 8: 70.8, 10: 72.6, 12: 76.9, 14: 85.1, 16: 94.7, 20: 95.0, 25: 108.0, 31: 123.3, 38: 146.1},
 256: {4: 116.6, 5: 112.1, 6: 110.1, 8: 112.9, 10: 114.3, 12: 117.9, 14: 122.6, 16: 130.6, 20: 132.9, 25: 150.5, 31: 175.6, 38: 179.8},
 512: {4: 236.9, 5: 221.5, 6: 212.7, 8: 207.1, 10: 207.2, 12: 208.4, 14: 212.1, 16: 217.8, 20: 225.5, 25: 241.9, 31: 257.6, 38: 277.5},
}
strips, slots = 129, 888
best = None
for c in np.arange(0.5, 4.01, 0.25):
    X, Y = [], []
    for ncol, d in meas.items():
        for Mx, t in d.items():
            X.append(makespan(uniform(ncol, Mx), strips, slots, c)); Y.append(t)
    X, Y = np.array(X), np.array(Y)
    A = np.vstack([X, np.ones_like(X)]).T
    (ts, t0), res, *_ = np.linalg.lstsq(A, Y, rcond=None)
    err = np.sqrt(np.mean((A @ [ts, t0] - Y) ** 2))
    print(f'c={c:.2f} step={ts:.3f} us t0={t0:.2f} rms={err:.2f}')
    if best is None or err < best[0]: best = (err, c, ts, t0)
print('best', best)

# ---- tapered (guided) chunk sequences: next width = clamp(remaining / (W f)), W = slots / strips concurrent chunk columns
def guided(ncol, W, f, lo, hi):
    out, r = [], ncol
    while r > 0:
        w = int(max(lo, min(hi, math.ceil(r / (W * f)))))
        w = min(w, r)
        out.append(w); r -= w
    return out
err, c, ts, t0 = best
for ncol in (124, 128, 256, 512, 1024):
    u = {Mx: makespan(uniform(ncol, Mx), strips, slots, c) * ts + t0 for Mx in (4, 5, 6, 8, 10, 12, 14, 16)}
    bu = min(u, key=u.get)
    line = f'ncol={ncol}: best uniform Mx={bu} {u[bu]:.1f} us |'
    for f in (1.0, 1.5, 2.0, 3.0):
        for hi in (8, 12, 16, 24):
            w = guided(ncol, slots / strips, f, 2, hi)
            line += f' f={f} hi={hi}: {makespan(w, strips, slots, c) * ts + t0:.1f} ({len(w)})'
        line += ' |'
    print(line)
