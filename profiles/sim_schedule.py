"""Offline model of SIMT scheduling policies for the 4-lanes-per-ray closest-hit kernel.

Input: per-ray step sequences of the exact traversal dumped by oracle_trace_steps (N = inner-node visit,
L = leaf visit with primitive tests, F = leaf visit whose box test fails, absorbed by the pop loop).
Output: warp-instructions per ray under each policy, given per-step instruction costs.
Design aid only; not used by tests or the product.
"""
import sys
import numpy as np

z = np.load(sys.argv[1] if len(sys.argv) > 1 else "/tmp/steps.npz")
buf, off = z["buf"], z["off"]
rays = []
for i in range(len(off) - 1):
    s = bytes(buf[off[i]:off[i + 1]]).replace(b"F", b"")
    rays.append(s)
CN, CL, CO, CSW = 60, 220, 40, 40
G = 8


def run_groups(policy, T=0, nwarps=64):
    """8 ray slots per warp with replacement."""
    per = len(rays) // nwarps
    total = 0
    for w in range(nwarps):
        queue = rays[w * per:(w + 1) * per]
        qi = 0
        slots = [None] * G  # (seq, pos)
        cost = 0
        while True:
            for g in range(G):
                if slots[g] is None and qi < len(queue):
                    while qi < len(queue) and len(queue[qi]) == 0:
                        qi += 1
                    if qi < len(queue):
                        slots[g] = [queue[qi], 0]
                        qi += 1
            if all(s is None for s in slots):
                break
            cost += CO
            state = lambda s: None if s is None else s[0][s[1]:s[1] + 1]
            if policy == "unified":
                anyN = any(state(s) == b"N" for s in slots)
                anyL = any(state(s) == b"L" for s in slots)
                cost += (CN if anyN else 0) + (CL if anyL else 0) - CO + 10
                for g in range(G):
                    if slots[g] is not None:
                        slots[g][1] += 1
                        if slots[g][1] >= len(slots[g][0]):
                            slots[g] = None
                continue
            # phase policy: node loop while more than T groups are in node state (or no leaf pending at all)
            while True:
                nN = sum(state(s) == b"N" for s in slots)
                nL = sum(state(s) == b"L" for s in slots)
                if nN == 0 or (nN <= T and nL > 0):
                    break
                cost += CN
                for g in range(G):
                    if state(slots[g]) == b"N":
                        slots[g][1] += 1
                        if slots[g][1] >= len(slots[g][0]):
                            slots[g] = None
            if any(state(s) == b"L" for s in slots):
                cost += CL
                for g in range(G):
                    if state(slots[g]) == b"L":
                        slots[g][1] += 1
                        if slots[g][1] >= len(slots[g][0]):
                            slots[g] = None
        total += cost
    return total / (per * nwarps)


def run_pool(P, nwarps=64):
    """P ray states per warp in shared memory; each iteration binds 8 of them to the 8 lane groups."""
    per = len(rays) // nwarps
    total = 0
    for w in range(nwarps):
        queue = [r for r in rays[w * per:(w + 1) * per] if len(r)]
        qi = 0
        pool = []
        cost = 0
        while True:
            while len(pool) < P and qi < len(queue):
                pool.append([queue[qi], 0]); qi += 1
            if not pool:
                break
            Ns = [s for s in pool if s[0][s[1]:s[1] + 1] == b"N"]
            Ls = [s for s in pool if s[0][s[1]:s[1] + 1] == b"L"]
            if len(Ls) >= G or not Ns:
                pick, c = Ls[:G], CL
            else:
                pick, c = Ns[:G], CN
            cost += c + CSW
            for s in pick:
                s[1] += 1
            pool = [s for s in pool if s[1] < len(s[0])]
        total += cost
    return total / (per * nwarps)


print(f"costs: node {CN}, leaf {CL}, outer {CO}, state switch {CSW}; mean steps/ray {np.mean([len(r) for r in rays]):.1f}")
ideal = np.mean([r.count(b'N') * CN + r.count(b'L') * CL for r in rays]) / G
print(f"ideal (perfect packing)      : {ideal:8.1f} warp-instr/ray")
for T in (0, 1, 2, 3, 4, 5, 6):
    print(f"phases, leave node loop at <={T}: {run_groups('phase', T):8.1f}")
print(f"unified                      : {run_groups('unified'):8.1f}")
for P in (12, 16, 24, 32):
    print(f"pool of {P:2d} rays             : {run_pool(P):8.1f}")
