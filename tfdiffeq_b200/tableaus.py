"""Butcher tableaus of the adaptive Runge-Kutta methods, as data for ``libb2ode``.

Row of SURVEY.md 8(a1): the reference keeps these as ``_ButcherTableau(alpha, beta, c_sol, c_error)`` python
lists plus a per-solver ``C_MID`` list (``tfdiffeq/rk_common.py:5``; dopri5 ``tfdiffeq/dopri5.py:11-36``,
dopri8 ``tfdiffeq/dopri8.py:12-79``, bosh3 ``tfdiffeq/bosh3.py:10-21``, tsit5 ``tfdiffeq/tsit5.py:10-30``,
adaptive heun ``tfdiffeq/adaptive_huen.py:11-25``).  Here a tableau is an immutable record that also
carries what the reference scatters over its solver classes: the order used for the initial step, the
order used by the controller, the dense-output flavour and the controller flavour.

Rationals are kept as exact ``Fraction`` s and converted once (``float(Fraction(a, b))`` is the correctly
rounded quotient, identical to python's ``a / b``); the embedded-error weights are the *float* difference
``float(b_j) - float(bhat_j)``, because that is what the reference's ``35 / 384 - 1951 / 21600`` evaluates to.
"""
from fractions import Fraction as Fr
from typing import NamedTuple, Optional, Tuple


class ButcherTableau(NamedTuple):
    name: str
    alpha: Tuple[float, ...]                 # s-1 nodes
    beta: Tuple[Tuple[float, ...], ...]      # s-1 rows, row i has i+1 weights
    c_sol: Tuple[float, ...]                 # s weights
    c_error: Tuple[float, ...]               # s weights of the embedded error estimate
    c_mid: Optional[Tuple[float, ...]]       # s weights of y(t0 + dt/2); None -> k-based dense output
    init_order: int                          # order handed to the initial-step heuristic
    ctrl_order: int                          # order in the step-size controller exponent
    controller: str = "reference"            # "reference" (misc.py:267-287) or "tsit5" (tsit5.py:53-62)

    @property
    def n_k(self):
        return len(self.c_sol)

    @property
    def fsal(self):
        """rk_common.py:54: the last stage input already is y1."""
        return self.c_sol[-1] == 0 and tuple(self.c_sol[:-1]) == tuple(self.beta[-1])


def _q(*pairs):
    return tuple(float(Fr(a, b)) if b else 0.0 for a, b in pairs)


def _diff(sol, hat):
    return tuple(s - h for s, h in zip(sol, hat))


# ---- Dormand-Prince 5(4) ------------------------------------------------------------------------
_dp_b = _q((35, 384), (0, 1), (500, 1113), (125, 192), (-2187, 6784), (11, 84))
_dp_bhat = _q((1951, 21600), (0, 1), (22642, 50085), (451, 720), (-12231, 42400), (649, 6300), (1, 60))
DOPRI5 = ButcherTableau(
    name="dopri5",
    alpha=_q((1, 5), (3, 10), (4, 5), (8, 9), (1, 1), (1, 1)),
    beta=(_q((1, 5)),
          _q((3, 40), (9, 40)),
          _q((44, 45), (-56, 15), (32, 9)),
          _q((19372, 6561), (-25360, 2187), (64448, 6561), (-212, 729)),
          _q((9017, 3168), (-355, 33), (46732, 5247), (49, 176), (-5103, 18656)),
          _dp_b),
    c_sol=_dp_b + (0.0,),
    c_error=_diff(_dp_b + (0.0,), _dp_bhat),
    # Shampine's mid-point weights, each halved exactly
    c_mid=tuple(v / 2 for v in _q((6025192743, 30085553152), (0, 1), (51252292925, 65400821598),
                                  (-2691868925, 45128329728), (187940372067, 1594534317056),
                                  (-1776094331, 19743644256), (11237099, 235043384))),
    init_order=4, ctrl_order=5)

# ---- Tsitouras 5(4), decimal literals as published ----------------------------------------------
_ts_b = (0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774)
_ts_bhat = (0.001780011052226, 0.000816434459657, -0.007880878010262, 0.144711007173263, -0.582357165452555,
            0.458082105929187)
TSIT5 = ButcherTableau(
    name="tsit5",
    alpha=(0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0),
    beta=((0.161,),
          (-0.008480655492357, 0.3354806554923570),
          (2.897153057105494, -6.359448489975075, 4.362295432869581),
          (5.32586482843925895, -11.74888356406283, 7.495539342889836, -0.09249506636175525),
          (5.86145544294642038, -12.92096931784711, 8.159367898576159, -0.071584973281401006, -0.02826905039406838),
          _ts_b),
    c_sol=_ts_b + (0.0,),
    c_error=_diff(_ts_b, _ts_bhat) + (-1.0 / 66.0,),
    c_mid=None, init_order=4, ctrl_order=5, controller="tsit5")

# ---- Bogacki-Shampine 3(2) ----------------------------------------------------------------------
_bs_b = (2.0 / 9.0, 1.0 / 3.0, 4.0 / 9.0)
_bs_err = (_bs_b[0] - 7.0 / 24.0, _bs_b[1] - 1.0 / 4.0, _bs_b[2] - 1.0 / 3.0, -1.0 / 8.0)
# as written in the reference: `1. / .2` and `3. / .4` (5.0 and 7.5 instead of 1/2 and 3/4), SURVEY App. A-6.
# Kept as the default because a drop-in must return what the reference returns.
BOSH3 = ButcherTableau(
    name="bosh3", alpha=(1.0 / 0.2, 3.0 / 4.0, 1.0), beta=((0.5,), (0.0, 3.0 / 0.4), _bs_b),
    c_sol=_bs_b + (0.0,), c_error=_bs_err, c_mid=(0.0, 0.5, 0.0, 0.0), init_order=2, ctrl_order=3)
# the textbook method, opt-in via odeint(..., method='bosh3', options={'textbook_tableau': True})
BOSH3_TEXTBOOK = BOSH3._replace(name="bosh3_textbook", alpha=(0.5, 0.75, 1.0), beta=((0.5,), (0.0, 0.75), _bs_b))

# ---- adaptive Heun 2(1); the reference hands order=5 to the controller (adaptive_huen.py:112) ----
ADAPTIVE_HEUN = ButcherTableau(
    name="adaptive_heun", alpha=(1.0,), beta=((1.0,),), c_sol=(0.5, 0.5), c_error=(0.5, -0.5), c_mid=(0.5, 0.0),
    init_order=1, ctrl_order=5)

# ---- Dormand-Prince 8(7), 13 stages -------------------------------------------------------------
_Z = (0, 1)
_d8_rows = (
    ((1, 18),),
    ((1, 48), (1, 16)),
    ((1, 32), _Z, (3, 32)),
    ((5, 16), _Z, (-75, 64), (75, 64)),
    ((3, 80), _Z, _Z, (3, 16), (3, 20)),
    ((29443841, 614563906), _Z, _Z, (77736538, 692538347), (-28693883, 1125000000), (23124283, 1800000000)),
    ((16016141, 946692911), _Z, _Z, (61564180, 158732637), (22789713, 633445777), (545815736, 2771057229),
     (-180193667, 1043307555)),
    ((39632708, 573591083), _Z, _Z, (-433636366, 683701615), (-421739975, 2616292301), (100302831, 723423059),
     (790204164, 839813087), (800635310, 3783071287)),
    ((246121993, 1340847787), _Z, _Z, (-37695042795, 15268766246), (-309121744, 1061227803),
     (-12992083, 490766935), (6005943493, 2108947869), (393006217, 1396673457), (123872331, 1001029789)),
    ((-1028468189, 846180014), _Z, _Z, (8478235783, 508512852), (1311729495, 1432422823),
     (-10304129995, 1701304382), (-48777925059, 3047939560), (15336726248, 1032824649),
     (-45442868181, 3398467696), (3065993473, 597172653)),
    ((185892177, 718116043), _Z, _Z, (-3185094517, 667107341), (-477755414, 1098053517), (-703635378, 230739211),
     (5731566787, 1027545527), (5232866602, 850066563), (-4093664535, 808688257), (3962137247, 1805957418),
     (65686358, 487910083)),
    ((403863854, 491063109), _Z, _Z, (-5068492393, 434740067), (-411421997, 543043805), (652783627, 914296604),
     (11173962825, 925320556), (-13158990841, 6184727034), (3936647629, 1978049680), (-160528059, 685178525),
     (248638103, 1413531060), _Z),
)
_d8_b = _q((14005451, 335480064), _Z, _Z, _Z, _Z, (-59238493, 1068277825), (181606767, 758867731),
           (561292985, 797845732), (-1041891430, 1371343529), (760417239, 1151165299), (118820643, 751138087),
           (-528747749, 2220607170), (1, 4))
_d8_bhat = _q((13451932, 455176623), _Z, _Z, _Z, _Z, (-808719846, 976000145), (1757004468, 5645159321),
              (656045339, 265891186), (-3867574721, 1518517206), (465885868, 322736535), (53011238, 667516719),
              (2, 45), _Z)

# dense-output weight polynomials of the 8(7) pair evaluated at theta = 1/2 (coefficients of theta^5..theta^0)
_D8_MID_POLY = {
    0: (-6.3448349392860401388, 22.1396504998094068976, -30.0610568289666450593, 19.9990069333683970610,
        -6.6910181737837595697, 1.0),
    5: (-39.6107919852202505218, 116.4422149550342161651, -121.4999627731334642623, 52.2273532792945524050,
        -7.6142658045872677172, None),
    6: (20.3761213808791436958, -67.1451318825957197185, 83.1721004639847717481, -46.8919164181093621583,
        10.7281392630428866124, None),
    7: (7.3347098826795362023, -16.5672243527496524646, 9.5724507555993664382, -0.1890893225010595467,
        0.5526637063753648783, None),
    8: (32.8801774352459155182, -89.9916014847245016028, 87.8406057677205645007, -35.7075975946222072821,
        4.2186562625665153803, None),
    9: (-10.1588990526426760954, 22.6237489648532849093, -17.4152107770762969005, 6.2736448083240352160,
        -0.6627209125361597559, None),
    10: (-12.5401268098782561200, 32.2362340167355370113, -28.5903289514790976966, 10.3160881272450748458,
         -1.2636789001135462218, None),
    11: (29.5553001484516038033, -82.1020315488359848644, 81.6630950584341412934, -34.7650769866611817349,
         5.4106037898590422230, None),
    12: (-41.7923486424390588923, 116.2662185791119533462, -114.9375291377009418170, 47.7457971078225540396,
         -7.0321379067945741781, None),
    13: (20.3006925822100825485, -53.9020777466385396792, 50.2558364226176017553, -19.0082099341608028453,
         2.3537586759714983486, None),
}


def _d8_mid():
    theta = 0.5
    out = [0.0] * 14
    for j, poly in _D8_MID_POLY.items():
        acc = None
        for power, coef in zip((5, 4, 3, 2, 1, 0), poly):
            if coef is None:
                continue
            term = coef * (theta ** power) if power else coef
            acc = term if acc is None else acc + term          # same left-to-right float sum as dopri8.py:52-72
        out[j] = acc / (1 / theta)
    return tuple(out)


DOPRI8 = ButcherTableau(
    name="dopri8",
    alpha=_q((1, 18), (1, 12), (1, 8), (5, 16), (3, 8), (59, 400), (93, 200), (5490023248, 9719169821), (13, 20),
             (1201146811, 1299019798), (1, 1), (1, 1), (1, 1)),
    beta=tuple(_q(*row) for row in _d8_rows) + (_d8_b,),
    c_sol=_d8_b + (0.0,),
    c_error=_diff(_d8_b, _d8_bhat) + (0.0,),
    c_mid=_d8_mid(), init_order=7, ctrl_order=8)

TABLEAUS = {t.name: t for t in (DOPRI5, TSIT5, BOSH3, BOSH3_TEXTBOOK, ADAPTIVE_HEUN, DOPRI8)}
