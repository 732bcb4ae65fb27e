"""Numerical check (numpy, CPU) of the tap-pairing rule for an N = 128 weight-gradient MMA.

A quadrant of such an MMA accumulates  S(a, b) = sum_q X[q + a] (x) dY[q + b]  over the pixels q of every tile of the image, with
TMA's zero fill outside the image for BOTH operands.  It equals the tap t = a - b of the 3x3 weight gradient
    dW[t] = sum_p X[p + t] (x) dY[p]      (zero padding of X)
if and only if, per axis, a = 0 or b = 0.  The script verifies the rule for all 81 (a, b) pairs on random images and then the
three-MMA cover of the nine taps + bias quoted in DESIGN.md.
"""
import itertools
import numpy as np

rng = np.random.default_rng(0)
H, W, CI, CO = 6, 5, 3, 4
X = rng.standard_normal((H, W, CI))
DY = rng.standard_normal((H, W, CO))


def at(T, y, x):
    return T[y, x] if 0 <= y < H and 0 <= x < W else np.zeros(T.shape[2])


def true_tap(t):
    return sum(np.outer(at(X, y + t[0], x + t[1]), DY[y, x]) for y in range(H) for x in range(W))


def quadrant(a, b):
    return sum(np.outer(at(X, y + a[0], x + a[1]), at(DY, y + b[0], x + b[1])) for y in range(H) for x in range(W))


S = [(y, x) for y in (-1, 0, 1) for x in (-1, 0, 1)]
rule_ok = lambda a, b: all(a[i] == 0 or b[i] == 0 for i in (0, 1))
n_valid = 0
for a, b in itertools.product(S, S):
    t = (a[0] - b[0], a[1] - b[1])
    exact = max(abs(t[0]), abs(t[1])) <= 1 and np.allclose(quadrant(a, b), true_tap(t), atol=1e-12)
    assert exact == rule_ok(a, b), (a, b, exact)
    n_valid += exact
print("rule 'per axis a = 0 or b = 0' <=> quadrant equals tap a - b: verified for all 81 shift pairs (%d valid)" % n_valid)

cover = [(((-1, -1), (-1, 0)), ((0, -1), (0, 0))), (((0, -1), (0, 0)), ((-1, 0), (0, 0))), (((0, 1), "ones"), ((-1, 0), (0, 0)))]
got = {}
for apair, bpair in cover:
    for a in apair:
        for b in bpair:
            if a == "ones":
                if b == (0, 0):
                    got["bias"] = sum(DY[y, x] for y in range(H) for x in range(W))
            elif rule_ok(a, b):
                got[(a[0] - b[0], a[1] - b[1])] = quadrant(a, b)
assert set(got) == set(S) | {"bias"}
for t in S:
    assert np.allclose(got[t], true_tap(t), atol=1e-12)
assert np.allclose(got["bias"], DY.sum((0, 1)))
print("three M = 128, N = 128 MMAs per k-step cover the nine taps and the bias: 10 of 12 quadrants used, all equal to the true gradient")
