"""ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's odeint() Runge-Kutta hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import this
module, and only as the checker / the timed CPU baseline -- never as the product path.  The product
(``tfdiffeq_b200``) never imports it and fails loudly when its CUDA library is missing.

Parity status: **pinned**.  ``oracle/make_golden.py`` runs the UNMODIFIED reference source
(``/root/reference/tfdiffeq/*.py`` over ``oracle/tf_shim.py``, TensorFlow itself is not installable
here) and stores its outputs in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this
restatement against every one of those vectors (values <= 1e-9 relative in fp64, identical accepted /
rejected / NFE counts) and against the reference tests' own analytic solutions.  Bit-level parity
with real TensorFlow kernels (``add_n`` order, ``pow`` last ulp) is *not* pinned by any reference
test (their tolerances are 1e-4 .. 1e-5) -- SURVEY.md section 8c.

Every function cites the reference ``file:line`` (relative to /root/reference) that it restates.
Arrays may be numpy arrays (the checker) or torch CPU tensors (``bench.py``'s multi-threaded CPU
baseline: op-for-op eager dispatch, the reference's own execution model); scalars are always numpy
scalars / python floats so the dtype of every scalar operation is explicit.
"""
import collections
import math
from fractions import Fraction
import warnings

import numpy as np

# --------------------------------------------------------------------------------------------------
# Butcher tableaus, restated from the reference files (values are the reference's, typos included).
# --------------------------------------------------------------------------------------------------


class Tableau(object):
    def __init__(self, name, alpha, beta, c_sol, c_error, c_mid, init_order, ctrl_order):
        self.name, self.alpha, self.beta = name, alpha, beta
        self.c_sol, self.c_error, self.c_mid = c_sol, c_error, c_mid
        self.init_order, self.ctrl_order = init_order, ctrl_order

    @property
    def fsal(self):
        # rk_common.py:54
        return self.c_sol[-1] == 0 and (self.c_sol[:-1] == self.beta[-1])


# tfdiffeq/dopri5.py:11-36 ; init order 4 (:74), controller order 5 (:68,:118)
DOPRI5 = Tableau(
    "dopri5",
    alpha=[1 / 5, 3 / 10, 4 / 5, 8 / 9, 1., 1.],
    beta=[[1 / 5],
          [3 / 40, 9 / 40],
          [44 / 45, -56 / 15, 32 / 9],
          [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
          [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
          [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]],
    c_sol=[35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0],
    c_error=[35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
             -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1. / 60.],
    c_mid=[6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
           187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2],
    init_order=4, ctrl_order=5)

# tfdiffeq/tsit5.py:10-30 ; init order 4 (:92), controller order 5 (:88)
TSIT5 = Tableau(
    "tsit5",
    alpha=[0.161, 0.327, 0.9, 0.9800255409045097, 1., 1.],
    beta=[[0.161],
          [-0.008480655492357, 0.3354806554923570],
          [2.897153057105494, -6.359448489975075, 4.362295432869581],
          [5.32586482843925895, -11.74888356406283, 7.495539342889836, -0.09249506636175525],
          [5.86145544294642038, -12.92096931784711, 8.159367898576159, -0.071584973281401006,
           -0.02826905039406838],
          [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081,
           2.324710524099774]],
    c_sol=[0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081,
           2.324710524099774, 0.],
    c_error=[0.09646076681806523 - 0.001780011052226, 0.01 - 0.000816434459657,
             0.4798896504144996 - (-0.007880878010262), 1.379008574103742 - 0.144711007173263,
             -3.290069515436081 - (-0.582357165452555), 2.324710524099774 - 0.458082105929187, -1. / 66.],
    c_mid=None, init_order=4, ctrl_order=5)

# tfdiffeq/bosh3.py:10-21 AS WRITTEN (typos `1./.2`, `3./.4`); init 2 (:56), ctrl 3 (:96)
BOSH3 = Tableau(
    "bosh3",
    alpha=[1. / .2, 3. / 4., 1.],
    beta=[[1. / 2.], [0., 3. / .4], [2. / 9., 1. / 3., 4. / 9.]],
    c_sol=[2. / 9., 1. / 3., 4. / 9., 0.],
    c_error=[2. / 9. - 7. / 24., 1. / 3. - 1. / 4., 4. / 9. - 1. / 3., -1. / 8.],
    c_mid=[0., 0.5, 0., 0.], init_order=2, ctrl_order=3)

# the textbook Bogacki-Shampine tableau (what bosh3.py evidently meant); opt-in, not reference behaviour
BOSH3_TEXTBOOK = Tableau(
    "bosh3_textbook",
    alpha=[1. / 2., 3. / 4., 1.],
    beta=[[1. / 2.], [0., 3. / 4.], [2. / 9., 1. / 3., 4. / 9.]],
    c_sol=BOSH3.c_sol, c_error=BOSH3.c_error, c_mid=BOSH3.c_mid, init_order=2, ctrl_order=3)

# tfdiffeq/adaptive_huen.py:11-25 ; init order 1 (:70), controller order 5 (:112, sic)
ADAPTIVE_HEUN = Tableau(
    "adaptive_heun", alpha=[1.], beta=[[1.]], c_sol=[0.5, 0.5], c_error=[0.5, -0.5], c_mid=[0.5, 0.],
    init_order=1, ctrl_order=5)


def _dopri8():
    # tfdiffeq/dopri8.py:12-79 ; init order 7 (:125), controller order 8 (:166)
    A = [1 / 18, 1 / 12, 1 / 8, 5 / 16, 3 / 8, 59 / 400, 93 / 200, 5490023248 / 9719169821, 13 / 20,
         1201146811 / 1299019798, 1, 1, 1]
    B = [
        [1 / 18],
        [1 / 48, 1 / 16],
        [1 / 32, 0, 3 / 32],
        [5 / 16, 0, -75 / 64, 75 / 64],
        [3 / 80, 0, 0, 3 / 16, 3 / 20],
        [29443841 / 614563906, 0, 0, 77736538 / 692538347, -28693883 / 1125000000, 23124283 / 1800000000],
        [16016141 / 946692911, 0, 0, 61564180 / 158732637, 22789713 / 633445777, 545815736 / 2771057229,
         -180193667 / 1043307555],
        [39632708 / 573591083, 0, 0, -433636366 / 683701615, -421739975 / 2616292301, 100302831 / 723423059,
         790204164 / 839813087, 800635310 / 3783071287],
        [246121993 / 1340847787, 0, 0, -37695042795 / 15268766246, -309121744 / 1061227803,
         -12992083 / 490766935, 6005943493 / 2108947869, 393006217 / 1396673457, 123872331 / 1001029789],
        [-1028468189 / 846180014, 0, 0, 8478235783 / 508512852, 1311729495 / 1432422823,
         -10304129995 / 1701304382, -48777925059 / 3047939560, 15336726248 / 1032824649,
         -45442868181 / 3398467696, 3065993473 / 597172653],
        [185892177 / 718116043, 0, 0, -3185094517 / 667107341, -477755414 / 1098053517, -703635378 / 230739211,
         5731566787 / 1027545527, 5232866602 / 850066563, -4093664535 / 808688257, 3962137247 / 1805957418,
         65686358 / 487910083],
        [403863854 / 491063109, 0, 0, -5068492393 / 434740067, -411421997 / 543043805, 652783627 / 914296604,
         11173962825 / 925320556, -13158990841 / 6184727034, 3936647629 / 1978049680, -160528059 / 685178525,
         248638103 / 1413531060, 0],
        [14005451 / 335480064, 0, 0, 0, 0, -59238493 / 1068277825, 181606767 / 758867731,
         561292985 / 797845732, -1041891430 / 1371343529, 760417239 / 1151165299, 118820643 / 751138087,
         -528747749 / 2220607170, 1 / 4]]
    C_sol = [14005451 / 335480064, 0, 0, 0, 0, -59238493 / 1068277825, 181606767 / 758867731,
             561292985 / 797845732, -1041891430 / 1371343529, 760417239 / 1151165299, 118820643 / 751138087,
             -528747749 / 2220607170, 1 / 4, 0]
    C_err = [14005451 / 335480064 - 13451932 / 455176623, 0, 0, 0, 0,
             -59238493 / 1068277825 - -808719846 / 976000145,
             181606767 / 758867731 - 1757004468 / 5645159321, 561292985 / 797845732 - 656045339 / 265891186,
             -1041891430 / 1371343529 - -3867574721 / 1518517206,
             760417239 / 1151165299 - 465885868 / 322736535, 118820643 / 751138087 - 53011238 / 667516719,
             -528747749 / 2220607170 - 2 / 45, 1 / 4, 0]
    h = 1 / 2
    polys = {  # dopri8.py:52-72: (h^5, h^4, h^3, h^2, h, const) coefficient rows of the dense-output weights
        0: (-6.3448349392860401388, 22.1396504998094068976, -30.0610568289666450593, 19.9990069333683970610,
            -6.6910181737837595697, 1.0),
        5: (-39.6107919852202505218, 116.4422149550342161651, -121.4999627731334642623,
            52.2273532792945524050, -7.6142658045872677172, 0.0),
        6: (20.3761213808791436958, -67.1451318825957197185, 83.1721004639847717481, -46.8919164181093621583,
            10.7281392630428866124, 0.0),
        7: (7.3347098826795362023, -16.5672243527496524646, 9.5724507555993664382, -0.1890893225010595467,
            0.5526637063753648783, 0.0),
        8: (32.8801774352459155182, -89.9916014847245016028, 87.8406057677205645007, -35.7075975946222072821,
            4.2186562625665153803, 0.0),
        9: (-10.1588990526426760954, 22.6237489648532849093, -17.4152107770762969005, 6.2736448083240352160,
            -0.6627209125361597559, 0.0),
        10: (-12.5401268098782561200, 32.2362340167355370113, -28.5903289514790976966, 10.3160881272450748458,
             -1.2636789001135462218, 0.0),
        11: (29.5553001484516038033, -82.1020315488359848644, 81.6630950584341412934, -34.7650769866611817349,
             5.4106037898590422230, 0.0),
        12: (-41.7923486424390588923, 116.2662185791119533462, -114.9375291377009418170,
             47.7457971078225540396, -7.0321379067945741781, 0.0),
        13: (20.3006925822100825485, -53.9020777466385396792, 50.2558364226176017553, -19.0082099341608028453,
             2.3537586759714983486, 0.0),
    }
    C_mid = np.zeros(14)
    for idx, (p5, p4, p3, p2, p1, p0) in polys.items():
        # identical expression order to dopri8.py:52-72
        if idx == 0:
            C_mid[idx] = (p5 * (h ** 5) + p4 * (h ** 4) + p3 * (h ** 3) + p2 * (h ** 2) + p1 * h + p0) / (1 / h)
        else:
            C_mid[idx] = (p5 * (h ** 5) + p4 * (h ** 4) + p3 * (h ** 3) + p2 * (h ** 2) + p1 * h) / (1 / h)
    return Tableau("dopri8", A, B, C_sol, C_err, C_mid.tolist(), init_order=7, ctrl_order=8)


DOPRI8 = _dopri8()

TABLEAUS = {"dopri5": DOPRI5, "tsit5": TSIT5, "bosh3": BOSH3, "bosh3_textbook": BOSH3_TEXTBOOK,
            "dopri8": DOPRI8, "adaptive_heun": ADAPTIVE_HEUN}

# --------------------------------------------------------------------------------------------------
# tiny array-backend helpers (numpy arrays or torch CPU tensors; scalars are numpy / python scalars)
# --------------------------------------------------------------------------------------------------


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _np_dtype(x):
    if _is_torch(x):
        return {"torch.float32": np.dtype(np.float32), "torch.float64": np.dtype(np.float64)}[str(x.dtype)]
    return x.dtype


def _absmax(x):
    """global max |x| as a numpy scalar of x's dtype (NaN propagates)."""
    dt = _np_dtype(x)
    if _is_torch(x):
        return dt.type(abs(x).max().item()) if x.numel() else dt.type(-np.inf)
    return dt.type(np.max(np.abs(x))) if x.size else dt.type(-np.inf)


def _mean(x):
    dt = _np_dtype(x)
    return dt.type(x.mean().item()) if _is_torch(x) else dt.type(np.mean(x))


def _l2(x):
    """tf.norm(x): sqrt(sum(x*x)) in x's dtype."""
    dt = _np_dtype(x)
    if _is_torch(x):
        return dt.type((x * x).sum().sqrt().item())
    return dt.type(np.sqrt(np.sum(x * x)))


def _numel(x):
    return x.numel() if _is_torch(x) else x.size


def _nonfinite(x):
    if _is_torch(x):
        import torch
        return bool((~torch.isfinite(x)).any())
    return bool(np.any(~np.isfinite(x)))


def _stack(xs):
    if _is_torch(xs[0]):
        import torch
        return torch.stack(list(xs))
    return np.stack(list(xs))


def _f(s):
    """numpy scalar -> python float (exact), so `scalar * array` keeps the array dtype in both backends."""
    return float(s)


# --------------------------------------------------------------------------------------------------
# hot-path arithmetic
# --------------------------------------------------------------------------------------------------


def scaled_dot_product(scale, xs, ys):
    """misc.py:118-121: add_n([(scale * x) * y ...]) summed left to right; `scale` is a state-dtype scalar.

    The reference's zero-skip test never fires (y is a Tensor, misc.py:114-121), so zero coefficients
    are still multiplied -- kept here so NaN propagation matches."""
    out = None
    for x, y in zip(xs, ys):
        term = _f(scale * type(scale)(x)) * y
        out = term if out is None else out + term
    return out


def dot_product(xs, ys):
    """misc.py:124-126: python sum() from 0 of x*y."""
    out = 0
    for x, y in zip(xs, ys):
        out = out + (x * y if not isinstance(x, np.generic) else _f(x) * y)
    return out


def runge_kutta_step(func, y0, f0, t0, dt, tableau):
    """rk_common.py:22-61.  y0/f0: tuples of arrays; t0/dt: float64 scalars (cast to state dtype :45-46)."""
    sd = _np_dtype(y0[0]).type
    t0 = sd(t0)
    dt = sd(dt)
    k = [[f] for f in f0]
    yi = None
    for alpha_i, beta_i in zip(tableau.alpha, tableau.beta):
        ti = t0 + sd(alpha_i) * dt
        yi = tuple(y0_ + scaled_dot_product(dt, beta_i, k_) for y0_, k_ in zip(y0, k))
        for k_, f_ in zip(k, func(ti, yi)):
            k_.append(f_)
    if not tableau.fsal:
        yi = tuple(y0_ + scaled_dot_product(dt, tableau.c_sol, k_) for y0_, k_ in zip(y0, k))
    y1 = yi
    f1 = tuple(k_[-1] for k_ in k)
    y1_error = tuple(scaled_dot_product(dt, tableau.c_error, k_) for k_ in k)
    return y1, f1, y1_error, k


def compute_error_ratio(error_estimate, rtol, atol, y0, y1):
    """misc.py:250-264: tol is a GLOBAL scalar per component (reduce_max over a python list, :257)."""
    out = []
    for err, atol_, rtol_, y0_, y1_ in zip(error_estimate, atol, rtol, y0, y1):
        sd = _np_dtype(y0_).type
        m = _nanmax([_absmax(y0_), _absmax(y1_)])
        tol = sd(atol_) + sd(rtol_) * m
        ratio = err / _f(tol)
        out.append(_mean(ratio * ratio))
    return tuple(out)


def _nanmax(vals):
    vals = list(vals)
    if any(np.isnan(v) for v in vals):
        return type(vals[0])(np.nan)
    return max(vals)


def optimal_step_size(last_step, mean_error_ratio, safety=0.9, ifactor=10.0, dfactor=0.2, order=5):
    """misc.py:267-287 (I-controller; exponent rounded through float32, :281-282)."""
    m = _nanmax(mean_error_ratio)
    last_step = np.float64(last_step)
    if m == 0:
        return last_step * np.float64(ifactor)
    if m < 1:
        dfactor = 1.0
    error_ratio = np.float64(np.sqrt(m))                 # sqrt in the state dtype, then cast to float64
    exponent = np.float64(np.float32(1. / order))
    with np.errstate(all="ignore"):
        cand = error_ratio ** exponent / np.float64(safety)
        inner = cand if np.isnan(cand) else min(cand, np.float64(1. / dfactor))
        factor = inner if np.isnan(inner) else max(np.float64(1. / ifactor), inner)
    return last_step / factor


def optimal_step_size_tsit5(last_step, mean_error_ratio, safety=0.9, ifactor=10.0, dfactor=0.2, order=5):
    """tsit5.py:53-62 (no square root; exponent is an exact float64 1/order)."""
    last_step = np.float64(last_step)
    m = mean_error_ratio
    if m == 0:
        return last_step * np.float64(ifactor)
    if m < 1:
        dfactor = 1.0
    error_ratio = np.float64(m)
    exponent = np.float64(1. / order)
    cand = error_ratio ** exponent / np.float64(safety)
    if np.isnan(cand):
        return last_step / cand
    factor = max(np.float64(1. / ifactor), min(cand, np.float64(1. / dfactor)))
    return last_step / factor


def norm_rms(x):
    """misc.py:170-175 for a single tensor: tf.norm(x) / numel(x)**0.5, all in x's dtype."""
    sd = _np_dtype(x).type
    return _l2(x) / (sd(_numel(x)) ** sd(0.5))


def select_initial_step(fun, t0, y0, order, rtol, atol, f0=None):
    """misc.py:183-247 (Hairer II.4).  Returns a state-dtype scalar; caller casts to float64."""
    sd = _np_dtype(y0[0]).type
    t0 = sd(t0)
    if f0 is None:
        f0 = fun(t0, y0)
    count = len(y0)
    rtol = list(rtol) if np.iterable(rtol) else [rtol] * count
    atol = list(atol) if np.iterable(atol) else [atol] * count
    scale = tuple(atol_ + abs(y0_) * rtol_ for y0_, atol_, rtol_ in zip(y0, atol, rtol))
    d0 = tuple(norm_rms(y0_ / scale_) for y0_, scale_ in zip(y0, scale))
    d1 = tuple(norm_rms(f0_ / scale_) for f0_, scale_ in zip(f0, scale))
    with np.errstate(all="ignore"):
        if max(d0) < 1e-5 or max(d1) < 1e-5:
            h0 = sd(1e-6)
        else:
            h0 = sd(0.01) * max(d0_ / d1_ for d0_, d1_ in zip(d0, d1))
        y1 = tuple(y0_ + _f(h0) * f0_ for y0_, f0_ in zip(y0, f0))
        f1 = fun(t0 + h0, y1)
        d2 = tuple(norm_rms((f1_ - f0_) / scale_) / h0 for f1_, f0_, scale_ in zip(f1, f0, scale))
        if max(d1) <= 1e-15 and max(d2) <= 1e-15:
            h1 = max(sd(1e-6), h0 * sd(1e-3))
        else:
            h1 = (sd(0.01) / max(d1 + d2)) ** sd(1. / float(order + 1))
        return min(sd(100) * h0, h1)


def interp_fit_rk(y0, y1, k, dt, tableau):
    """dopri5.py:39-45 / bosh3.py:24-30 / dopri8.py:82-87 / adaptive_huen.py:28-34."""
    sd = _np_dtype(y0[0]).type
    dt = sd(dt)
    y_mid = tuple(y0_ + scaled_dot_product(dt, tableau.c_mid, k_) for y0_, k_ in zip(y0, k))
    f0 = tuple(k_[0] for k_ in k)
    f1 = tuple(k_[-1] for k_ in k)
    # interp.py:22-36 multiplies python ints by the dt *tensor*: -2*dt etc. evaluated in the state dtype
    return _interp_fit_typed(y0, y1, y_mid, f0, f1, dt)


def _interp_fit_typed(y0, y1, y_mid, f0, f1, dt):
    sd = type(dt)
    m2, p2, p5, m3, m4 = _f(sd(-2) * dt), _f(sd(2) * dt), _f(sd(5) * dt), _f(sd(-3) * dt), _f(sd(-4) * dt)
    d1 = _f(dt)

    def dp(cs, ts):
        out = 0
        for c_, t_ in zip(cs, ts):
            out = out + c_ * t_
        return out
    a = tuple(dp([m2, p2, -8, -8, 16], [f0_, f1_, y0_, y1_, ym_]) for f0_, f1_, y0_, y1_, ym_ in
              zip(f0, f1, y0, y1, y_mid))
    b = tuple(dp([p5, m3, 18, 14, -32], [f0_, f1_, y0_, y1_, ym_]) for f0_, f1_, y0_, y1_, ym_ in
              zip(f0, f1, y0, y1, y_mid))
    c = tuple(dp([m4, d1, -11, -5, 16], [f0_, f1_, y0_, y1_, ym_]) for f0_, f1_, y0_, y1_, ym_ in
              zip(f0, f1, y0, y1, y_mid))
    d = tuple(d1 * f0_ for f0_ in f0)
    return [a, b, c, d, y0]


def interp_evaluate(coefficients, t0, t1, t):
    """interp.py:39-67: x = (t - t0) / (t1 - t0) in the STATE dtype (:55-60); a x^4 + b x^3 + c x^2 + d x + e
    with explicit powers, summed left to right from 0."""
    sd = _np_dtype(coefficients[0][0]).type
    t0, t1, t = sd(t0), sd(t1), sd(t)
    assert (t0 <= t) & (t <= t1), 'invalid interpolation, fails `t0 <= t <= t1`: {}, {}, {}'.format(t0, t, t1)
    with np.errstate(all="ignore"):
        x = sd((t - t0) / (t1 - t0))
    xs = [sd(1), x]
    for _ in range(2, len(coefficients)):
        xs.append(xs[-1] * x)
    rx = [_f(v) for v in reversed(xs)]
    out = []
    for coeffs in zip(*coefficients):
        acc = 0
        for c_, x_ in zip(coeffs, rx):
            acc = acc + c_ * x_
        out.append(acc)
    return tuple(out)


def interp_eval_tsit5(t0, t1, k, eval_t):
    """tsit5.py:33-50, bugs included: `y0 = k_[0]` is f0, not the state (:47).  t0,t1,eval_t float64."""
    dt = np.float64(t1) - np.float64(t0)
    t = (np.float64(eval_t) - np.float64(t0)) / dt
    b1 = -1.0530884977290216 * t * (t - 1.3299890189751412) * (t ** 2 - 1.4364028541716351 * t + 0.7139816917074209)
    b2 = 0.1017 * t ** 2 * (t ** 2 - 2.1966568338249754 * t + 1.2949852507374631)
    b3 = 2.490627285651252793 * t ** 2 * (t ** 2 - 2.38535645472061657 * t + 1.57803468208092486)
    b4 = -16.54810288924490272 * (t - 1.21712927295533244) * (t - 0.61620406037800089) * t ** 2
    b5 = 47.37952196281928122 * (t - 1.203071208372362603) * (t - 0.658047292653547382) * t ** 2
    b6 = -34.87065786149660974 * (t - 1.2) * (t - 0.666666666666666667) * t ** 2
    b7 = 2.5 * (t - 1) * (t - 0.6) * t ** 2
    coeff = [b1, b2, b3, b4, b5, b6, b7]
    y0 = tuple(k_[0] for k_ in k)
    out = []
    for y0_, k_ in zip(y0, k):
        acc = None
        for c_, kk in zip(coeff, k_):
            term = _f(dt * c_) * kk
            acc = term if acc is None else acc + term
        out.append(y0_ + acc)
    return tuple(out)


# --------------------------------------------------------------------------------------------------
# drivers
# --------------------------------------------------------------------------------------------------


class Stats(object):
    def __init__(self):
        self.n_acc = 0
        self.n_rej = 0
        self.nfe = 0
        self.dt_trace = []      # dt of every attempted step
        self.acc_trace = []     # accept flag of every attempted step


def _tf_f64(v):
    """misc.py:137-144 `_convert_to_tensor(a, dtype=tf.float64)`: python floats become float32 first."""
    if isinstance(v, float):
        return np.float64(np.float32(v))
    return np.float64(v)


def _listify(v, n):
    return list(v) if np.iterable(v) else [v] * n


class AdaptiveRK(object):
    """solvers.py:27-35 + dopri5.py:50-121 (same skeleton in bosh3.py, dopri8.py, adaptive_huen.py)."""

    def __init__(self, func, y0, rtol, atol, tableau, first_step=None, safety=0.9, ifactor=10.0, dfactor=0.2,
                 max_num_steps=2 ** 31 - 1, stats=None, **unused):
        if unused:
            warnings.warn('{}: Unexpected arguments {}'.format(type(self).__name__, unused))
        self.func, self.y0, self.tab = func, y0, tableau
        self.rtol, self.atol = _listify(rtol, len(y0)), _listify(atol, len(y0))
        self.first_step = first_step
        # dopri5.py:62-64: _convert_to_tensor(python float, dtype=float64) goes THROUGH float32 first
        # (tf.convert_to_tensor(0.9) is float32), so safety = 0.8999999761581421, dfactor = 0.20000000298023224
        self.safety, self.ifactor, self.dfactor = _tf_f64(safety), _tf_f64(ifactor), _tf_f64(dfactor)
        self.max_num_steps = max_num_steps
        self.stats = stats if stats is not None else Stats()

    def _f(self, t, y):
        self.stats.nfe += 1
        return self.func(t, y)

    def before_integrate(self, t):
        sd = _np_dtype(self.y0[0]).type
        f0 = self._f(sd(t[0]), self.y0)
        if self.first_step is None:
            fs = select_initial_step(self._f, t[0], self.y0, self.tab.init_order, self.rtol[0], self.atol[0], f0=f0)
            first_step = np.float64(fs)
        else:
            first_step = _tf_f64(self.first_step)
        # _RungeKuttaState(y1, f1, t0, t1, dt, interp_coeff)
        self.state = (self.y0, f0, np.float64(t[0]), np.float64(t[0]), first_step, [self.y0] * 5)

    def advance(self, next_t):
        n_steps = 0
        while next_t > self.state[3]:
            assert n_steps < self.max_num_steps, 'max_num_steps exceeded ({}>={})'.format(n_steps, self.max_num_steps)
            self.state = self._step(self.state)
            n_steps += 1
        return interp_evaluate(self.state[5], self.state[2], self.state[3], next_t)

    def _step(self, st):
        y0, f0, _, t0, dt, coeff = st          # NB the field called t1 is unpacked as t0 (dopri5.py:93)
        dt = np.float64(dt)
        assert t0 + dt > t0, 'underflow in dt {}'.format(dt)
        for y0_ in y0:
            assert not _nonfinite(y0_), 'non-finite values in state `y`: {}'.format(y0_)
        y1, f1, y1_err, k = runge_kutta_step(self._f, y0, f0, t0, dt, self.tab)
        with np.errstate(all="ignore"):
            msr = compute_error_ratio(y1_err, self.rtol, self.atol, y0, y1)
        accept = all(bool(m <= 1) for m in msr)
        self.stats.dt_trace.append(float(dt))
        self.stats.acc_trace.append(accept)
        if accept:
            self.stats.n_acc += 1
        else:
            self.stats.n_rej += 1
        y_next = y1 if accept else y0
        f_next = f1 if accept else f0
        t_next = t0 + dt if accept else t0
        coeff = interp_fit_rk(y0, y1, k, dt, self.tab) if accept else coeff
        dt_next = optimal_step_size(dt, msr, self.safety, self.ifactor, self.dfactor, self.tab.ctrl_order)
        return (y_next, f_next, t0, t_next, dt_next, coeff)

    def integrate(self, t):
        t = np.asarray(t, dtype=np.float64)          # solvers.py:30
        assert np.all(t[1:] > t[:-1]), 't must be strictly increasing or decrasing'
        solution = [self.y0]
        self.before_integrate(t)
        for i in range(1, t.shape[0]):
            solution.append(self.advance(t[i]))
        return tuple(_stack(s) for s in zip(*solution))


class Tsit5(AdaptiveRK):
    """tsit5.py:65-151: pooled error over all components, controller without sqrt, k-based (buggy) dense output."""

    def __init__(self, func, y0, rtol, atol, **kw):
        AdaptiveRK.__init__(self, func, y0, rtol, atol, TSIT5, **kw)
        self.rtol, self.atol = rtol, atol          # scalars, not listified (tsit5.py:81-82)

    def before_integrate(self, t):
        if self.first_step is None:
            fs = select_initial_step(self._f, t[0], self.y0, 4, self.rtol, self.atol)   # recomputes f0 (:92)
            first_step = np.float64(fs)
        else:
            first_step = _tf_f64(self.first_step)
        f0 = self._f(np.float64(t[0]), self.y0)                                        # :98, t not cast
        self.state = (self.y0, f0, np.float64(t[0]), np.float64(t[0]), first_step, [self.y0] * 7)

    def advance(self, next_t):
        n_steps = 0
        while next_t > self.state[3]:
            assert n_steps < self.max_num_steps, 'max_num_steps exceeded ({}>={})'.format(n_steps, self.max_num_steps)
            self.state = self._step(self.state)
            n_steps += 1
        return interp_eval_tsit5(self.state[2], self.state[3], self.state[5], next_t)

    def _step(self, st):
        y0, f0, _, t0, dt, _ = st
        assert t0 + dt > t0, 'underflow in dt {}'.format(dt)
        for y0_ in y0:
            assert not _nonfinite(y0_), 'non-finite values in state `y`: {}'.format(y0_)
        y1, f1, y1_err, k = runge_kutta_step(self._f, y0, f0, t0, dt, TSIT5)
        sd = _np_dtype(y0[0]).type
        total, count = sd(0), 0
        with np.errstate(all="ignore"):
            for err, y0_, y1_ in zip(y1_err, y0, y1):
                tol = sd(self.atol) + sd(self.rtol) * _nanmax([_absmax(y0_), _absmax(y1_)])
                r = err / _f(tol)
                sq = r * r
                total = total + sd(sq.sum().item() if _is_torch(sq) else np.sum(sq))
                count += _numel(sq)
            mean_error_ratio = total / sd(count)
        accept = bool(mean_error_ratio <= 1.)
        self.stats.dt_trace.append(float(dt))
        self.stats.acc_trace.append(accept)
        if accept:
            self.stats.n_acc += 1
        else:
            self.stats.n_rej += 1
        y_next = y1 if accept else y0
        f_next = f1 if accept else f0
        t_next = t0 + dt if accept else t0
        dt_next = optimal_step_size_tsit5(dt, mean_error_ratio, self.safety, self.ifactor, self.dfactor, 5)
        k_next = k if accept else st[5]
        return (y_next, f_next, t0, t_next, dt_next, k_next)


class FixedGrid(object):
    """solvers.py:39-115 + fixed_grid.py + rk_common.py:73-81.  method in {euler, midpoint, heun, rk4}."""

    def __init__(self, func, y0, method, step_size=None, grid_constructor=None, eps=0.0, stats=None, **unused):
        unused.pop('rtol', None)
        unused.pop('atol', None)
        if unused:
            warnings.warn('{}: Unexpected arguments {}'.format(type(self).__name__, unused))
        self.func, self.y0, self.method, self.eps = func, y0, method, eps
        self.stats = stats if stats is not None else Stats()
        if step_size is not None and grid_constructor is None:
            self.grid_constructor = self._grid_from_step_size(step_size)
        elif grid_constructor is None:
            self.grid_constructor = lambda f, y0, t: t
        else:
            raise ValueError("step_size and grid_constructor are exclusive arguments.")

    @staticmethod
    def _grid_from_step_size(step_size):
        # solvers.py:58-71 is broken under TF2 (tf.ceil); this is the evident intent
        def ctor(func, y0, t):
            sd = t.dtype.type
            start, end = t[0], t[-1]
            niters = int(math.ceil(float((end - start) / sd(step_size) + 1)))
            g = np.arange(0, niters).astype(t.dtype) * sd(step_size) + start
            if g[-1] > t[-1]:
                g[-1] = t[-1]
            return g
        return ctor

    def _f(self, t, y):
        self.stats.nfe += 1
        return self.func(t, y)

    def step_func(self, t, dt, y):
        sd = type(dt)
        eps = sd(self.eps)
        d = _f(dt)
        if self.method == "euler":                                   # fixed_grid.py:6-7
            return tuple(d * f_ for f_ in self._f(t + eps, y))
        if self.method == "midpoint":                                # fixed_grid.py:16-18
            y_mid = tuple(y_ + f_ * d / 2 for y_, f_ in zip(y, self._f(t + eps, y)))
            return tuple(d * f_ for f_ in self._f(t + dt / sd(2), y_mid))
        if self.method == "heun":                                    # fixed_grid.py:28-32
            f_outs = self._f(t + eps, y)
            hat = tuple(y_ + d * f_ for y_, f_ in zip(y, f_outs))
            f1 = self._f(t + dt, hat)
            return tuple(_f(dt / sd(2.)) * (a_ + b_) for a_, b_ in zip(f_outs, f1))
        if self.method == "rk4":                                     # rk_common.py:73-81 via fixed_grid.py:41-42
            t = t + eps
            k1 = self._f(t, y)
            k2 = self._f(t + dt / sd(3), tuple(y_ + d * k1_ / 3 for y_, k1_ in zip(y, k1)))
            k3 = self._f(t + dt * sd(2) / sd(3), tuple(y_ + d * (k1_ / -3 + k2_) for y_, k1_, k2_ in zip(y, k1, k2)))
            k4 = self._f(t + dt, tuple(y_ + d * (k1_ - k2_ + k3_) for y_, k1_, k2_, k3_ in zip(y, k1, k2, k3)))
            return tuple((k1_ + 3 * k2_ + 3 * k3_ + k4_) * _f(dt / sd(8))
                         for k1_, k2_, k3_, k4_ in zip(k1, k2, k3, k4))
        raise KeyError(self.method)

    def integrate(self, t):
        sd = _np_dtype(self.y0[0])
        t = np.asarray(t).astype(sd)                                  # solvers.py:84
        assert np.all(t[1:] > t[:-1]), 't must be strictly increasing or decrasing'
        grid = self.grid_constructor(self.func, self.y0, t)
        assert grid[0] == t[0] and grid[-1] == t[-1]
        solution = [self.y0]
        j = 1
        y0 = self.y0
        for t0, t1 in zip(grid[:-1], grid[1:]):
            dy = self.step_func(t0, t1 - t0, y0)
            y1 = tuple(y0_ + dy_ for y0_, dy_ in zip(y0, dy))
            while j < t.shape[0] and t1 >= t[j]:
                solution.append(self._linear_interp(t0, t1, y0, y1, t[j]))
                j += 1
            y0 = y1
        return tuple(_stack(s) for s in zip(*solution))

    @staticmethod
    def _linear_interp(t0, t1, y0, y1, t):
        # solvers.py:106-115
        if t == t0:
            return y0
        if t == t1:
            return y1
        slope = tuple((y1_ - y0_) / _f(t1 - t0) for y0_, y1_ in zip(y0, y1))
        return tuple(y0_ + s_ * _f(t - t0) for y0_, s_ in zip(y0, slope))


ADAPTIVE = {"dopri5": DOPRI5, "dopri8": DOPRI8, "bosh3": BOSH3, "adaptive_heun": ADAPTIVE_HEUN}
FIXED = {"euler": "euler", "midpoint": "midpoint", "rk4": "rk4", "huen": "heun", "heun": "heun"}


# --------------------------------------------------------------------------------------------------
# multistep solvers: fixed_adams.py and adams.py
# --------------------------------------------------------------------------------------------------


def _poly_mul(a, b):
    out = [Fraction(0)] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            out[i + j] += x * y
    return out


def _adams_weights(k, shift):
    """Weights of the k-point Adams formula as exact rationals: integral over [0, 1] of the Lagrange basis on the
    nodes u = shift, shift - 1, ..., shift - k + 1 (shift = 0: Bashforth, nodes t_n, t_n-1, ...; shift = 1: Moulton,
    nodes t_n+1, t_n, ...).  Returns (integer numerators, common divisor) like the tables of fixed_adams.py:7-160."""
    nodes = [Fraction(shift - j) for j in range(k)]
    w = []
    for j in range(k):
        poly, den = [Fraction(1)], Fraction(1)
        for i in range(k):
            if i != j:
                poly = _poly_mul(poly, [-nodes[i], Fraction(1)])
                den *= nodes[j] - nodes[i]
        w.append(sum(c / (p + 1) for p, c in enumerate(poly)) / den)
    div = 1
    for x in w:
        div = div * x.denominator // math.gcd(div, x.denominator)
    return [int(x * div) for x in w], div


def has_converged(y0, y1, rtol, atol):
    """misc.py:129-134: every element within atol + rtol * max(|y0|, |y1|) (element-wise tolerance here)."""
    for a, b in zip(y0, y1):
        if _is_torch(a):
            import torch
            tol = atol + rtol * torch.maximum(a.abs(), b.abs())
            ok = bool(((a - b).abs() < tol).all())
        else:
            tol = atol + rtol * np.maximum(np.abs(a), np.abs(b))
            ok = bool(np.all(np.abs(a - b) < tol))
        if not ok:
            return False
    return True


class FixedAdams(FixedGrid):
    """fixed_adams.py:168-212.  implicit=True: 'fixed_adams' (Bashforth predictor + Moulton corrector by functional
    iteration); implicit=False: 'explicit_adams'.  The first steps (fewer than 3 stored derivatives) are 3/8-rule
    RK4 steps reusing the stored derivative as k1 (:188-191)."""

    MIN_ORDER, MAX_ORDER, MAX_ITERS = 4, 12, 4

    def __init__(self, func, y0, rtol=1e-3, atol=1e-4, implicit=True, max_iters=4, max_order=12, stats=None, **kw):
        FixedGrid.__init__(self, func, y0, "adams", stats=stats, **kw)
        self.rtol, self.atol, self.implicit, self.max_iters = rtol, atol, implicit, max_iters
        self.max_order = int(min(max_order, self.MAX_ORDER))
        self.prev_f = collections.deque(maxlen=self.max_order - 1)
        self.prev_t = None

    def _update_history(self, t, f):
        if self.prev_t is None or self.prev_t != t:                  # fixed_adams.py:182-185
            self.prev_f.appendleft(f)
            self.prev_t = t

    def step_func(self, t, dt, y):
        sd = type(dt)
        d = _f(dt)
        self._update_history(t, self._f(t, y))
        order = min(len(self.prev_f), self.max_order - 1)
        if order < self.MIN_ORDER - 1:                               # :190-193, rk_common.py:73-81 with k1 given
            k1 = self.prev_f[0]
            k2 = self._f(t + dt / sd(3), tuple(y_ + d * k1_ / 3 for y_, k1_ in zip(y, k1)))
            k3 = self._f(t + dt * sd(2) / sd(3), tuple(y_ + d * (k1_ / -3 + k2_) for y_, k1_, k2_ in zip(y, k1, k2)))
            k4 = self._f(t + dt, tuple(y_ + d * (k1_ - k2_ + k3_) for y_, k1_, k2_, k3_ in zip(y, k1, k2, k3)))
            return tuple((k1_ + 3 * k2_ + 3 * k3_ + k4_) * _f(dt / sd(8)) for k1_, k2_, k3_, k4_ in zip(k1, k2, k3, k4))
        ab, ab_div = _adams_weights(order, 0)
        dy = tuple(d * _sdp_py(1 / ab_div, ab, f_) for f_ in zip(*self.prev_f))          # :196-198
        if self.implicit:
            am, am_div = _adams_weights(order + 1, 1)
            delta = tuple(d * _sdp_py(1 / am_div, am[1:], f_) for f_ in zip(*self.prev_f))
            converged = False
            f = None
            for _ in range(self.max_iters):
                dy_old = dy
                f = self._f(t + dt, tuple(y_ + dy_ for y_, dy_ in zip(y, dy)))
                c0 = _f(dt * sd(am[0] / am_div))                      # dt (state dtype) * python float, then * f
                dy = tuple(c0 * f_ + delta_ for f_, delta_ in zip(f, delta))
                converged = has_converged(dy_old, dy, self.rtol, self.atol)
                if converged:
                    break
            if not converged:
                self.stats.not_converged = getattr(self.stats, "not_converged", 0) + 1
                self.prev_f.pop()                                     # :210 drops the OLDEST stored derivative
            self._update_history(t, f)                                # :211 no-op: prev_t == t already
        return dy


def _sdp_py(scale, xs, ys):
    """misc.py:118-121 with a python-float scale and integer coefficients: ((scale * x) * y) summed left to right."""
    out = None
    for x, y in zip(xs, ys):
        term = (scale * x) * y
        out = term if out is None else out + term
    return out


GAMMA_STAR = [1, -1 / 2, -1 / 12, -1 / 24, -19 / 720, -3 / 160, -863 / 60480, -275 / 24192, -33953 / 3628800,
              -0.00789255, -0.00678585, -0.00592406, -0.00523669, -0.0046775, -0.00421495, -0.0038269]   # adams.py:16-19


class VariableAdams(object):
    """adams.py:22-211: variable-coefficient Adams-Bashforth-Moulton (Hairer-Norsett-Wanner III.5), orders 1..12.

    Quirks kept: g is stored in a float32 Variable (:34, :51, :56); `first_step` is ignored (:112-115); the
    predictor uses order - 1 terms (:144-147); the state carried to the next step is the PREDICTOR p_next, not the
    corrected y_next (:211); a rejected step keeps the order (:172)."""

    MIN_ORDER, MAX_ORDER = 1, 12

    def __init__(self, func, y0, rtol, atol, implicit=True, first_step=None, max_order=12, safety=0.9, ifactor=10.0,
                 dfactor=0.2, stats=None, **unused):
        if unused:
            warnings.warn('{}: Unexpected arguments {}'.format('VariableCoefficientAdamsBashforth', unused))
        self.func, self.y0 = func, y0
        self.rtol, self.atol = _listify(rtol, len(y0)), _listify(atol, len(y0))
        self.max_order = int(max(self.MIN_ORDER, min(max_order, self.MAX_ORDER)))
        self.safety, self.ifactor, self.dfactor = _tf_f64(safety), _tf_f64(ifactor), _tf_f64(dfactor)
        self.stats = stats if stats is not None else Stats()

    def _f(self, t, y):
        self.stats.nfe += 1
        return self.func(t, y)

    def integrate(self, t):
        t = np.asarray(t, dtype=np.float64)                           # solvers.py:30
        assert np.all(t[1:] > t[:-1]), 't must be strictly increasing or decrasing'
        sd = _np_dtype(self.y0[0]).type
        f0 = self._f(sd(t[0]), self.y0)
        first_step = np.float64(select_initial_step(self._f, t[0], self.y0, 2, self.rtol[0], self.atol[0], f0=f0))
        self.y_n, self.prev_t, self.phi = self.y0, collections.deque([t[0]], maxlen=self.max_order + 1), [f0]
        self.n_prev_f = 1
        self.next_t, self.order = t[0] + first_step, 1
        solution = [self.y0]
        for i in range(1, len(t)):
            while t[i] > self.prev_t[0]:                              # adams.py:123-126
                self._step(t[i])
            assert t[i] == self.prev_t[0]
            solution.append(self.y_n)
        return tuple(_stack(s) for s in zip(*solution))

    def _g_and_explicit_phi(self, next_t, k):
        """adams.py:29-59."""
        prev_t, iphi = self.prev_t, self.phi
        curr_t = prev_t[0]
        dt = next_t - prev_t[0]
        g = np.zeros(k + 1, dtype=np.float32)
        g[0] = 1
        c = 1.0 / np.arange(1, k + 2).astype(np.float64)
        ephi = [iphi[0]]
        beta = np.float64(1.0)
        sd = _np_dtype(iphi[0][0]).type
        with np.errstate(all="ignore"):
            for j in range(1, k):
                beta = (next_t - prev_t[j - 1]) / (curr_t - prev_t[j]) * beta
                ephi.append(tuple(p_ * _f(sd(beta)) for p_ in iphi[j]))
                c = c[:-1] - c[1:] if j == 1 else c[:-1] - c[1:] * dt / (next_t - prev_t[j - 1])
                g[j] = np.float32(c[0])
            c = c[:-1] - c[1:] * dt / (next_t - prev_t[k - 1])
            g[k] = np.float32(c[0])
        return g, ephi

    @staticmethod
    def _implicit_phi(ephi, f_n, k):
        """adams.py:62-77."""
        k = min(len(ephi) + 1, k)
        out = [f_n]
        for j in range(1, k):
            out.append(tuple(a_ - b_ for a_, b_ in zip(out[j - 1], ephi[j - 1])))
        return out

    def _error(self, coef, phi_j, tol):
        """misc.py:250-264 with the tolerance given: mean((coef * phi / tol)**2) per component."""
        out = []
        for p_, tol_ in zip(phi_j, tol):
            ratio = (_f(coef) * p_) / _f(tol_)
            out.append(_mean(ratio * ratio))
        return tuple(out)

    def _step(self, final_t):
        """adams.py:128-211."""
        y0, order = self.y_n, self.order
        sd = _np_dtype(y0[0]).type
        next_t = min(self.next_t, final_t) if not np.isnan(self.next_t) else self.next_t
        dt = next_t - self.prev_t[0]
        dtc = sd(dt)
        g32, phi = self._g_and_explicit_phi(next_t, order)
        g = g32.astype(sd)
        m = max(1, order - 1)
        p_next = tuple(y0_ + scaled_dot_product(dtc, g[:m], phi_[:m]) for y0_, phi_ in zip(y0, tuple(zip(*phi))))
        next_f0 = self._f(sd(next_t), p_next)
        iphi_p = self._implicit_phi(phi, next_f0, order + 1)
        y_next = tuple(p_ + _f(dtc * g[order - 1]) * i_ for p_, i_ in zip(p_next, iphi_p[order - 1]))
        tol = tuple(sd(a_) + sd(r_) * _nanmax([_absmax(a0), _absmax(a1)])
                    for a_, r_, a0, a1 in zip(self.atol, self.rtol, y0, y_next))
        error_k = self._error(dtc * (g[order] - g[order - 1]), iphi_p[order], tol)
        accept = all(bool(e <= 1) for e in error_k)
        if not accept:
            self.stats.n_rej += 1
            dt_next = optimal_step_size(dt, error_k, self.safety, self.ifactor, self.dfactor, order=order)
            self.next_t = self.prev_t[0] + dt_next
            return
        self.stats.n_acc += 1
        next_f0 = self._f(sd(next_t), y_next)
        implicit_phi = self._implicit_phi(phi, next_f0, order + 2)
        next_order = order
        if len(self.prev_t) <= 4 or order < 3:
            next_order = min(order + 1, 3, self.max_order)
        else:
            e1 = self._error(dtc * (g[order - 1] - g[order - 2]), iphi_p[order - 1], tol)
            e2 = self._error(dtc * (g[order - 2] - g[order - 3]), iphi_p[order - 2], tol)
            if min(e1 + e2) < max(error_k):
                next_order = order - 1
            elif order < self.max_order:
                ep = self._error(dtc * sd(GAMMA_STAR[order]), iphi_p[order], tol)
                if max(ep) < max(error_k):
                    next_order = order + 1
        dt_next = dt if next_order > order else optimal_step_size(dt, error_k, self.safety, self.ifactor,
                                                                  self.dfactor, order=order + 1)
        self.prev_t.appendleft(next_t)
        self.y_n, self.phi, self.order = p_next, implicit_phi, next_order      # :211 the predictor is carried on
        self.next_t = next_t + dt_next


def odeint(func, y0, t, rtol=1e-7, atol=1e-9, method=None, options=None, stats=None):
    """odeint.py:28-81 + misc.py:290-329 (_check_inputs).  y0: array or tuple of arrays; t: 1-D array."""
    tensor_input = False
    if not isinstance(y0, tuple):
        tensor_input = True
        y0 = (y0,)
        base = func
        func = lambda t_, y_: (base(t_, y_[0]),)         # noqa: E731  misc.py:301-303
    t = np.asarray(t)
    if bool(np.all(t[1:] < t[:-1])):                      # misc.py:318-321 (a length-1 t counts as decreasing)
        t = -t
        rev = func
        func = lambda t_, y_: tuple(-f_ for f_ in rev(-t_, y_))   # noqa: E731
    if options is None:
        options = {}
    elif method is None:
        raise ValueError('cannot supply `options` without specifying `method`')
    if method is None:
        method = 'dopri5'
    if method in ADAPTIVE:
        solver = AdaptiveRK(func, y0, rtol, atol, ADAPTIVE[method], stats=stats, **options)
    elif method == "bosh3_textbook":
        solver = AdaptiveRK(func, y0, rtol, atol, BOSH3_TEXTBOOK, stats=stats, **options)
    elif method == "tsit5":
        solver = Tsit5(func, y0, rtol, atol, stats=stats, **options)
    elif method in FIXED:
        solver = FixedGrid(func, y0, FIXED[method], stats=stats, **options)
    elif method == "fixed_adams":
        solver = FixedAdams(func, y0, rtol=rtol, atol=atol, stats=stats, **options)
    elif method == "explicit_adams":
        solver = FixedAdams(func, y0, rtol=rtol, atol=atol, implicit=False, stats=stats, **options)
    elif method == "adams":
        solver = VariableAdams(func, y0, rtol, atol, stats=stats, **options)
    else:
        raise KeyError(method)
    sol = solver.integrate(t)
    return sol[0] if tensor_input else sol
