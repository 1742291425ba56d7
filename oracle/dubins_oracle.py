"""CPU restatement of the `dubins` package's shortest-path solver and sampler.  TEST INFRASTRUCTURE ONLY.

The reference calls a third-party dependency that is absent from /root/reference and from this image:

    /root/reference/paths_library.py:150-152   dubins.shortest_path(cp, goal, self.tr)  ->  path.sample_many(self.ss / 10.0)
    (same two calls at :203-205, :263-264, :325-328)

`dubins` is the PyPI Cython wrapper (pydubins, unpinned by the reference -- README.md:34 names no version; the last
release is 1.0.1) around Andrew Walker's C library "Dubins-Curves" (src/dubins.c).  Its published algorithm, restated
here function by function:

  dubins_shortest_path   normalise (d = D / rho, theta = mod2pi(atan2(dy, dx)), alpha = mod2pi(th0 - theta),
                         beta = mod2pi(th1 - theta)); evaluate the six words in the order of the library's enum
                         DubinsPathType { LSL, LSR, RSL, RSR, RLR, LRL }; keep a word only when its cost is STRICTLY
                         below the best so far -- so exact ties go to the earlier word of that order
  dubins_LSL ... _LRL    the closed forms of Shkel & Lumelsky, "Classification of the Dubins set" (2001)
  dubins_segment         end configuration of one L / S / R piece from (x, y, heading) in units of rho
  dubins_path_sample     configuration at arc length t: `if tprime < p1 ... else if tprime < p1 + p2 ... else`
  dubins_path_sample_many  `x = 0; while x < length: sample(x); x += stepSize` -- the end point is NOT appended; whether
                         a sample lands on it is decided by the accumulated floating-point sum
  pydubins sample_many   returns (list of (x, y, heading) tuples, list of arc lengths)

PARITY PINNING.  The library is not in the image, so this restatement cannot be checked against its binary.  It is
pinned instead by construction (tests/test_dubins.py): every word's three pieces are re-integrated independently
(rigid rotation about the turning-circle centre, and an RK4 integration of the unicycle ODE) and must end on the goal
pose to 1e-9; the chosen word's length must equal the minimum over a GEOMETRIC brute force -- tangent-line /
tangent-circle constructions, a derivation independent of the closed forms -- over all six word families on >= 10 000
random (pose, goal, rho) triples, including the rho = 0.05 of BASELINE config 1.
"""
import math

import numpy as np

TWO_PI = 2.0 * math.pi
WORDS = ('LSL', 'LSR', 'RSL', 'RSR', 'RLR', 'LRL')        # DubinsPathType enum order = tie-break order


def mod2pi(theta):
    """dubins.c: fmodr(theta, 2 pi) = x - y * floor(x / y)."""
    return theta - TWO_PI * math.floor(theta / TWO_PI)


def words(alpha, beta, d):
    """[(word, t, p, q) or None] in enum order (dubins.c: dubins_LSL ... dubins_LRL; None = EDUBNOPATH)."""
    sa, sb, ca, cb = math.sin(alpha), math.sin(beta), math.cos(alpha), math.cos(beta)
    c_ab = math.cos(alpha - beta)
    d_sq = d * d
    out = []
    # LSL
    p_sq = 2 + d_sq - (2 * c_ab) + (2 * d * (sa - sb))
    if p_sq >= 0:
        tmp1 = math.atan2(cb - ca, d + sa - sb)
        out.append(('LSL', mod2pi(tmp1 - alpha), math.sqrt(p_sq), mod2pi(beta - tmp1)))
    else:
        out.append(None)
    # LSR
    p_sq = -2 + d_sq + (2 * c_ab) + (2 * d * (sa + sb))
    if p_sq >= 0:
        p = math.sqrt(p_sq)
        tmp0 = math.atan2(-ca - cb, d + sa + sb) - math.atan2(-2.0, p)
        out.append(('LSR', mod2pi(tmp0 - alpha), p, mod2pi(tmp0 - mod2pi(beta))))
    else:
        out.append(None)
    # RSL
    p_sq = -2 + d_sq + (2 * c_ab) - (2 * d * (sa + sb))
    if p_sq >= 0:
        p = math.sqrt(p_sq)
        tmp0 = math.atan2(ca + cb, d - sa - sb) - math.atan2(2.0, p)
        out.append(('RSL', mod2pi(alpha - tmp0), p, mod2pi(beta - tmp0)))
    else:
        out.append(None)
    # RSR
    p_sq = 2 + d_sq - (2 * c_ab) + (2 * d * (sb - sa))
    if p_sq >= 0:
        tmp1 = math.atan2(ca - cb, d - sa + sb)
        out.append(('RSR', mod2pi(alpha - tmp1), math.sqrt(p_sq), mod2pi(tmp1 - beta)))
    else:
        out.append(None)
    # RLR
    tmp0 = (6. - d_sq + 2 * c_ab + 2 * d * (sa - sb)) / 8.
    if abs(tmp0) <= 1:
        phi = math.atan2(ca - cb, d - sa + sb)
        p = mod2pi((2 * math.pi) - math.acos(tmp0))
        t = mod2pi(alpha - phi + mod2pi(p / 2.))
        out.append(('RLR', t, p, mod2pi(alpha - beta - t + mod2pi(p))))
    else:
        out.append(None)
    # LRL
    tmp0 = (6. - d_sq + 2 * c_ab + 2 * d * (sb - sa)) / 8.
    if abs(tmp0) <= 1:
        phi = math.atan2(ca - cb, d + sa - sb)
        p = mod2pi(2 * math.pi - math.acos(tmp0))
        t = mod2pi(-alpha - phi + p / 2.)
        out.append(('LRL', t, p, mod2pi(mod2pi(beta) - alpha - t + mod2pi(p))))
    else:
        out.append(None)
    return out


def normalise(q0, q1, rho):
    dx, dy = q1[0] - q0[0], q1[1] - q0[1]
    D = math.sqrt(dx * dx + dy * dy)
    d = D / rho
    theta = mod2pi(math.atan2(dy, dx)) if d > 0 else 0.0
    return mod2pi(q0[2] - theta), mod2pi(q1[2] - theta), d


def shortest_path(q0, q1, rho):
    """dubins_shortest_path: (word, (t, p, q)) in units of rho; strict `<` over the enum order."""
    alpha, beta, d = normalise(q0, q1, rho)
    best, best_cost = None, float('inf')
    for w in words(alpha, beta, d):
        if w is None:
            continue
        cost = w[1] + w[2] + w[3]
        if cost < best_cost:
            best, best_cost = w, cost
    return best[0], best[1:]


def segment(t, qi, kind):
    """dubins_segment: configuration after travelling t along an L / S / R piece from qi (units of rho)."""
    st, ct = math.sin(qi[2]), math.cos(qi[2])
    if kind == 'L':
        qt = (+math.sin(qi[2] + t) - st, -math.cos(qi[2] + t) + ct, t)
    elif kind == 'R':
        qt = (-math.sin(qi[2] - t) + st, +math.cos(qi[2] - t) - ct, -t)
    else:
        qt = (ct * t, st * t, 0.0)
    return (qt[0] + qi[0], qt[1] + qi[1], qt[2] + qi[2])


def path_sample(q0, word, params, rho, t):
    """dubins_path_sample."""
    tprime = t / rho
    qi = (0.0, 0.0, q0[2])
    p1, p2 = params[0], params[1]
    q1 = segment(p1, qi, word[0])
    q2 = segment(p2, q1, word[1])
    if tprime < p1:
        q = segment(tprime, qi, word[0])
    elif tprime < (p1 + p2):
        q = segment(tprime - p1, q1, word[1])
    else:
        q = segment(tprime - p1 - p2, q2, word[2])
    return (q[0] * rho + q0[0], q[1] * rho + q0[1], mod2pi(q[2]))


def sample_many(q0, word, params, rho, step):
    """dubins_path_sample_many through pydubins: ([(x, y, heading)], [arc length])."""
    length = (params[0] + params[1] + params[2]) * rho
    x = 0.0
    qs, ts = [], []
    while x < length:
        qs.append(path_sample(q0, word, params, rho, x))
        ts.append(x)
        x += step
    return qs, ts


# ---------------------------------------------------------------------------------------------------------------------
# Independent checks (no closed form of the library is used below)

def _centre(q, side):
    """Centre of the unit turning circle left ('L') or right ('R') of configuration q."""
    s = 1.0 if side == 'L' else -1.0
    return np.array([q[0] - s * math.sin(q[2]), q[1] + s * math.cos(q[2])])


def integrate_rigid(q, word, params):
    """End configuration of the three pieces by rigid motions: an arc is a rotation about the circle centre."""
    x, y, th = q
    for kind, t in zip(word, params):
        if kind == 'S':
            x, y = x + t * math.cos(th), y + t * math.sin(th)
        else:
            c = _centre((x, y, th), kind)
            a = t if kind == 'L' else -t
            r = np.array([x, y]) - c
            ca, sa = math.cos(a), math.sin(a)
            x, y = c[0] + ca * r[0] - sa * r[1], c[1] + sa * r[0] + ca * r[1]
            th = th + a
    return x, y, th


def integrate_rk4(q, word, params, steps=4000):
    """End configuration by RK4 on the unicycle ODE x' = cos th, y' = sin th, th' = u (u = +1, 0, -1)."""
    state = np.array(q, dtype=float)
    for kind, t in zip(word, params):
        u = {'L': 1.0, 'S': 0.0, 'R': -1.0}[kind]
        h = t / steps

        def f(s):
            return np.array([math.cos(s[2]), math.sin(s[2]), u])
        for _ in range(steps):
            k1 = f(state)
            k2 = f(state + 0.5 * h * k1)
            k3 = f(state + 0.5 * h * k2)
            k4 = f(state + h * k3)
            state = state + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
    return tuple(state)


def geometric_candidates(q0n, q1n):
    """Every Dubins candidate between two configurations in units of rho, by tangent constructions:
    CSC words from the common tangent lines of the two turning circles, CCC words from the two circles tangent to both.
    Returns [(word, (t, p, q))]; each is verified by the caller (integration must end on q1n)."""
    th0, th1 = q0n[2], q1n[2]
    out = []
    for a, b in (('L', 'L'), ('R', 'R'), ('L', 'R'), ('R', 'L')):
        c0, c1 = _centre(q0n, a), _centre(q1n, b)
        v = c1 - c0
        ell = math.hypot(v[0], v[1])
        gamma = math.atan2(v[1], v[0])
        if a == b:
            if ell == 0.0:
                continue
            phi, p = gamma, ell                              # outer tangent: parallel to the line of centres
        else:
            if ell < 2.0:
                continue                                     # circles overlap: no inner tangent
            p = math.sqrt(max(ell * ell - 4.0, 0.0))
            # leaving an L circle at heading phi and entering an R circle: v = R(phi) (p, -2); R -> L: v = R(phi) (p, +2)
            phi = gamma + math.atan2(2.0, p) if a == 'L' else gamma - math.atan2(2.0, p)
        t = mod2pi(phi - th0) if a == 'L' else mod2pi(th0 - phi)
        q = mod2pi(th1 - phi) if b == 'L' else mod2pi(phi - th1)
        out.append((a + 'S' + b, (t, p, q)))
    for outer in ('R', 'L'):
        c0, c1 = _centre(q0n, outer), _centre(q1n, outer)
        v = c1 - c0
        ell = math.hypot(v[0], v[1])
        if ell > 4.0 or ell == 0.0:
            continue
        gamma = math.atan2(v[1], v[0])
        a = math.acos(min(1.0, ell / 4.0))
        for sign in (+1.0, -1.0):
            psi0 = gamma + sign * a                           # direction c0 -> middle centre
            m = c0 + 2.0 * np.array([math.cos(psi0), math.sin(psi0)])
            psi1 = math.atan2(m[1] - c1[1], m[0] - c1[0])     # direction c1 -> middle centre
            if outer == 'R':                                  # on an R circle the radius to the vehicle is (-sin, cos)(heading)
                phi1, phi2 = psi0 - math.pi / 2, psi1 - math.pi / 2
                t, p, q = mod2pi(th0 - phi1), mod2pi(phi2 - phi1), mod2pi(phi2 - th1)
                out.append(('RLR', (t, p, q)))
            else:
                phi1, phi2 = psi0 + math.pi / 2, psi1 + math.pi / 2
                t, p, q = mod2pi(phi1 - th0), mod2pi(phi1 - phi2), mod2pi(th1 - phi2)
                out.append(('LRL', (t, p, q)))
    return out


def pose_error(end, goal):
    dth = (end[2] - goal[2] + math.pi) % TWO_PI - math.pi
    return max(abs(end[0] - goal[0]), abs(end[1] - goal[1]), abs(dth))


def brute_force_shortest(q0, q1, rho, tol=1e-9):
    """Minimum total length (units of rho) over the verified geometric candidates, and the candidate list."""
    q0n = (0.0, 0.0, q0[2])
    q1n = ((q1[0] - q0[0]) / rho, (q1[1] - q0[1]) / rho, q1[2])
    scale = 1.0 + abs(q1n[0]) + abs(q1n[1])
    ok = []
    for word, prm in geometric_candidates(q0n, q1n):
        end = integrate_rigid(q0n, word, prm)
        if pose_error(end, q1n) <= tol * scale:
            ok.append((sum(prm), word, prm))
    ok.sort(key=lambda c: c[0])
    return ok
