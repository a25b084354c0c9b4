"""NumPy FP64 restatement of near-earth SGP4 as ``dsgp4`` computes it.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference calls
``dsgp4.initialize_tle`` / ``dsgp4.propagate`` / ``dsgp4.propagate_batch``
(/root/reference/kessler/model.py:581,592,536,541; util.py:570-576); dsgp4
itself (esa/dSGP4, un-pinned in /root/reference/setup.py:35) is not vendored,
so this file restates its published algorithm -- Vallado, Crawford, Hujsak,
Kelso, "Revisiting Spacetrack Report #3", AIAA 2006-6753, near-earth branch,
WGS-84 constants, opsmode 'i', floor-mod angle wrapping, Kepler loop <= 10
iterations to 1e-12 with a +-0.95 step clamp -- and is
pinned against /root/reference/docs/notebooks/trace.pickle through
tests/golden/trace_golden.json (tests/test_oracle_golden.py).

Layout of the per-TLE constant record (shared with oracle/csrc/oracle.c and
include/kessler_b200.h, ``KSL_C_*``): see ``CONST_NAMES``.
"""
import datetime
import math

import numpy as np

# WGS-84, as dsgp4.util.get_gravity_constants('wgs-84') (model.py:571); values
# equal the pickled _mu,_radiusearthkm,_xke,_tumin,_j2,_j3,_j4,_j3oj2.
MU = 398600.5
RE = 6378.137
XKE = 60.0 / math.sqrt(RE * RE * RE / MU)
TUMIN = 1.0 / XKE
J2 = 0.00108262998905
J3 = -0.00000253215306
J4 = -0.00000161098761
J3OJ2 = J3 / J2
X2O3 = 2.0 / 3.0
TWOPI = 2.0 * math.pi
VKMPERSEC = RE * XKE / 60.0

# error bits (bitmask; dsgp4 keeps only the last code in tle._error and
# Kessler never reads it -- only the deep-space raise is observable,
# model.py:580-599)
ERR_ECC = 1        # em >= 1 or em < -0.001          (Vallado error 1)
ERR_MEANMOTION = 2  # nm <= 0                        (Vallado error 2)
ERR_SEMILATUS = 4  # pl < 0                          (Vallado error 4)
ERR_DECAYED = 8    # mrt < 1                         (Vallado error 6)
ERR_DEEPSPACE = 16  # period >= 225 min: dsgp4.initialize_tle raises

CONST_NAMES = (
    "no_unkozai", "ecco", "inclo", "nodeo", "argpo", "mo", "bstar",
    "mdot", "argpdot", "nodedot", "nodecf", "cc1", "cc4", "cc5", "t2cof",
    "omgcof", "xmcof", "eta", "delmo", "sinmao", "d2", "d3", "d4",
    "t3cof", "t4cof", "t5cof", "aycof", "xlcof", "con41", "x1mth2",
    "x7thm1", "isimp", "cosio", "sinio", "aobase", "pad",
)
NCONST = len(CONST_NAMES)  # 36


def sgp4_init(no_kozai, ecco, inclo, nodeo, argpo, mo, bstar):
    """Vallado ``sgp4init`` (+ ``initl``), vectorised.  Inputs in rad/min and
    rad.  Returns (dict of arrays keyed by CONST_NAMES plus a/alta/altp, err)."""
    no_kozai, ecco, inclo, nodeo, argpo, mo, bstar = [
        np.asarray(x, dtype=np.float64) for x in (no_kozai, ecco, inclo, nodeo, argpo, mo, bstar)]
    with np.errstate(all="ignore"):
        eccsq = ecco * ecco
        omeosq = 1.0 - eccsq
        rteosq = np.sqrt(omeosq)
        cosio = np.cos(inclo)
        cosio2 = cosio * cosio
        ak = np.power(XKE / no_kozai, X2O3)
        d1 = 0.75 * J2 * (3.0 * cosio2 - 1.0) / (rteosq * omeosq)
        del_ = d1 / (ak * ak)
        adel = ak * (1.0 - del_ * del_ - del_ * (1.0 / 3.0 + 134.0 * del_ * del_ / 81.0))
        del_ = d1 / (adel * adel)
        no = no_kozai / (1.0 + del_)
        ao = np.power(XKE / no, X2O3)
        sinio = np.sin(inclo)
        po = ao * omeosq
        con42 = 1.0 - 5.0 * cosio2
        con41 = -con42 - cosio2 - cosio2
        posq = po * po
        rp = ao * (1.0 - ecco)

        isimp = rp < (220.0 / RE + 1.0)
        sfour = np.full_like(rp, 78.0 / RE + 1.0)
        qzms24 = np.full_like(rp, ((120.0 - 78.0) / RE) ** 4)
        perige = (rp - 1.0) * RE
        low = perige < 156.0
        sf_low = np.where(perige < 98.0, 20.0, perige - 78.0)
        qz_low = ((120.0 - sf_low) / RE) ** 4
        sfour = np.where(low, sf_low / RE + 1.0, sfour)
        qzms24 = np.where(low, qz_low, qzms24)

        pinvsq = 1.0 / posq
        tsi = 1.0 / (ao - sfour)
        eta = ao * ecco * tsi
        etasq = eta * eta
        eeta = ecco * eta
        psisq = np.abs(1.0 - etasq)
        coef = qzms24 * tsi ** 4
        coef1 = coef / psisq ** 3.5
        cc2 = coef1 * no * (ao * (1.0 + 1.5 * etasq + eeta * (4.0 + etasq))
                            + 0.375 * J2 * tsi / psisq * con41 * (8.0 + 3.0 * etasq * (8.0 + etasq)))
        cc1 = bstar * cc2
        bige = ecco > 1.0e-4
        safe_e = np.where(bige, ecco, 1.0)
        cc3 = np.where(bige, -2.0 * coef * tsi * J3OJ2 * no * sinio / safe_e, 0.0)
        x1mth2 = 1.0 - cosio2
        cc4 = 2.0 * no * coef1 * ao * omeosq * (
            eta * (2.0 + 0.5 * etasq) + ecco * (0.5 + 2.0 * etasq)
            - J2 * tsi / (ao * psisq) * (-3.0 * con41 * (1.0 - 2.0 * eeta + etasq * (1.5 - 0.5 * eeta))
                                         + 0.75 * x1mth2 * (2.0 * etasq - eeta * (1.0 + etasq)) * np.cos(2.0 * argpo)))
        cc5 = 2.0 * coef1 * ao * omeosq * (1.0 + 2.75 * (etasq + eeta) + eeta * etasq)
        cosio4 = cosio2 * cosio2
        temp1 = 1.5 * J2 * pinvsq * no
        temp2 = 0.5 * temp1 * J2 * pinvsq
        temp3 = -0.46875 * J4 * pinvsq * pinvsq * no
        mdot = no + 0.5 * temp1 * rteosq * con41 + 0.0625 * temp2 * rteosq * (13.0 - 78.0 * cosio2 + 137.0 * cosio4)
        argpdot = (-0.5 * temp1 * con42 + 0.0625 * temp2 * (7.0 - 114.0 * cosio2 + 395.0 * cosio4)
                   + temp3 * (3.0 - 36.0 * cosio2 + 49.0 * cosio4))
        xhdot1 = -temp1 * cosio
        nodedot = xhdot1 + (0.5 * temp2 * (4.0 - 19.0 * cosio2) + 2.0 * temp3 * (3.0 - 7.0 * cosio2)) * cosio
        omgcof = bstar * cc3 * np.cos(argpo)
        safe_eeta = np.where(bige, eeta, 1.0)
        xmcof = np.where(bige, -X2O3 * coef * bstar / safe_eeta, 0.0)
        nodecf = 3.5 * omeosq * xhdot1 * cc1
        t2cof = 1.5 * cc1
        den = np.where(np.abs(cosio + 1.0) > 1.5e-12, 1.0 + cosio, 1.5e-12)
        xlcof = -0.25 * J3OJ2 * sinio * (3.0 + 5.0 * cosio) / den
        aycof = -0.5 * J3OJ2 * sinio
        delmotemp = 1.0 + eta * np.cos(mo)
        delmo = delmotemp * delmotemp * delmotemp
        sinmao = np.sin(mo)
        x7thm1 = 7.0 * cosio2 - 1.0

        cc1sq = cc1 * cc1
        d2 = 4.0 * ao * tsi * cc1sq
        temp = d2 * tsi * cc1 / 3.0
        d3 = (17.0 * ao + sfour) * temp
        d4 = 0.5 * temp * ao * tsi * (221.0 * ao + 31.0 * sfour) * cc1
        t3cof = d2 + 2.0 * cc1sq
        t4cof = 0.25 * (3.0 * d3 + cc1 * (12.0 * d2 + 10.0 * cc1sq))
        t5cof = 0.2 * (3.0 * d4 + 12.0 * cc1 * d3 + 6.0 * d2 * d2 + 15.0 * cc1sq * (2.0 * d2 + cc1sq))
        zero = np.zeros_like(d2)
        d2, d3, d4, t3cof, t4cof, t5cof = [np.where(isimp, zero, x) for x in (d2, d3, d4, t3cof, t4cof, t5cof)]

        a = np.power(no * TUMIN, -2.0 / 3.0)
        err = np.where(TWOPI / no >= 225.0, ERR_DEEPSPACE, 0).astype(np.int32)
        # not-a-number / non-positive mean motion cannot be initialised either
        err = np.where(np.isfinite(no) & (no > 0.0), err, ERR_DEEPSPACE).astype(np.int32)

    c = dict(no_unkozai=no, ecco=ecco, inclo=inclo, nodeo=nodeo, argpo=argpo, mo=mo, bstar=bstar,
             mdot=mdot, argpdot=argpdot, nodedot=nodedot, nodecf=nodecf, cc1=cc1, cc4=cc4, cc5=cc5,
             t2cof=t2cof, omgcof=omgcof, xmcof=xmcof, eta=eta, delmo=delmo, sinmao=sinmao,
             d2=d2, d3=d3, d4=d4, t3cof=t3cof, t4cof=t4cof, t5cof=t5cof, aycof=aycof, xlcof=xlcof,
             con41=con41, x1mth2=x1mth2, x7thm1=x7thm1, isimp=isimp.astype(np.float64),
             cosio=cosio, sinio=sinio, aobase=ao, pad=np.zeros_like(ao),
             a=a, alta=a * (1.0 + ecco) - 1.0, altp=a * (1.0 - ecco) - 1.0)
    return c, err


def consts_to_array(c):
    """dict from sgp4_init -> SoA array [NCONST, n] in CONST_NAMES order."""
    n = np.asarray(c["no_unkozai"]).reshape(-1).shape[0]
    out = np.zeros((NCONST, n), dtype=np.float64)
    for i, k in enumerate(CONST_NAMES):
        out[i] = np.asarray(c[k], dtype=np.float64).reshape(-1)
    return out


def array_to_consts(arr):
    return {k: arr[i] for i, k in enumerate(CONST_NAMES)}


def sgp4_propagate(c, t, couple_kepler=True, return_mean=False):
    """SGP4 state at ``t`` minutes since epoch.  ``c`` entries broadcast against
    ``t`` (e.g. c[...] of shape [n,1] with t [n,m], or scalars with t [m]).
    Returns r [km] (...,3), v [km/s] (...,3), err bitmask (...).

    ``couple_kepler=True`` iterates every lane until the slowest has converged,
    as a vectorised dsgp4.propagate call does ([recall]);
    False stops each lane on its own (scalar Vallado semantics)."""
    t = np.asarray(t, dtype=np.float64)
    g = lambda k: np.asarray(c[k], dtype=np.float64)
    with np.errstate(all="ignore"):
        xmdf = g("mo") + g("mdot") * t
        argpdf = g("argpo") + g("argpdot") * t
        nodedf = g("nodeo") + g("nodedot") * t
        argpm = argpdf
        mm = xmdf
        t2 = t * t
        nodem = nodedf + g("nodecf") * t2
        tempa = 1.0 - g("cc1") * t
        tempe = g("bstar") * g("cc4") * t
        templ = g("t2cof") * t2

        full = g("isimp") == 0.0
        delomg = g("omgcof") * t
        delmtemp = 1.0 + g("eta") * np.cos(xmdf)
        delm = g("xmcof") * (delmtemp * delmtemp * delmtemp - g("delmo"))
        temp = delomg + delm
        mm_f = xmdf + temp
        argpm_f = argpdf - temp
        t3 = t2 * t
        t4 = t3 * t
        tempa_f = tempa - g("d2") * t2 - g("d3") * t3 - g("d4") * t4
        tempe_f = tempe + g("bstar") * g("cc5") * (np.sin(mm_f) - g("sinmao"))
        templ_f = templ + g("t3cof") * t3 + t4 * (g("t4cof") + t * g("t5cof"))
        mm = np.where(full, mm_f, mm)
        argpm = np.where(full, argpm_f, argpm)
        tempa = np.where(full, tempa_f, tempa)
        tempe = np.where(full, tempe_f, tempe)
        templ = np.where(full, templ_f, templ)

        no = g("no_unkozai")
        am = np.power(XKE / no, X2O3) * tempa * tempa
        nm = XKE / np.power(am, 1.5)
        em = g("ecco") - tempe
        err = np.zeros(np.broadcast(em, t).shape, dtype=np.int32)
        err |= np.where((em >= 1.0) | (em < -0.001), ERR_ECC, 0).astype(np.int32)
        err |= np.where(nm <= 0.0, ERR_MEANMOTION, 0).astype(np.int32)
        em = np.where(em < 1.0e-6, 1.0e-6, em)
        mm = mm + no * templ
        xlm = mm + argpm + nodem
        nodem = np.mod(nodem, TWOPI)
        argpm = np.mod(argpm, TWOPI)
        xlm = np.mod(xlm, TWOPI)
        mm = np.mod(xlm - argpm - nodem, TWOPI)
        mean = (am, em, nodem, argpm, mm, nm)

        sinim, cosim = g("sinio"), g("cosio")
        axnl = em * np.cos(argpm)
        temp = 1.0 / (am * (1.0 - em * em))
        aynl = em * np.sin(argpm) + temp * g("aycof")
        xl = mm + argpm + nodem + temp * g("xlcof") * axnl
        u = np.mod(xl - nodem, TWOPI)
        eo1 = u.copy()
        tem5 = np.full_like(eo1, 9999.9)
        sineo1 = np.zeros_like(eo1)
        coseo1 = np.ones_like(eo1)
        ktr = 1
        while ktr <= 10:
            active = np.abs(tem5) >= 1.0e-12
            if not np.any(active):
                break
            if couple_kepler:
                active = np.ones_like(active)
            s, co = np.sin(eo1), np.cos(eo1)
            d = 1.0 - co * axnl - s * aynl
            d = (u - aynl * co + axnl * s - eo1) / d
            d = np.where(d >= 0.95, 0.95, d)
            d = np.where(d <= -0.95, -0.95, d)
            sineo1 = np.where(active, s, sineo1)
            coseo1 = np.where(active, co, coseo1)
            tem5 = np.where(active, d, tem5)
            eo1 = np.where(active, eo1 + d, eo1)
            ktr += 1

        ecose = axnl * coseo1 + aynl * sineo1
        esine = axnl * sineo1 - aynl * coseo1
        el2 = axnl * axnl + aynl * aynl
        pl = am * (1.0 - el2)
        err |= np.where(pl < 0.0, ERR_SEMILATUS, 0).astype(np.int32)
        rl = am * (1.0 - ecose)
        rdotl = np.sqrt(am) * esine / rl
        rvdotl = np.sqrt(pl) / rl
        betal = np.sqrt(1.0 - el2)
        temp = esine / (1.0 + betal)
        sinu = am / rl * (sineo1 - aynl - axnl * temp)
        cosu = am / rl * (coseo1 - axnl + aynl * temp)
        su = np.arctan2(sinu, cosu)
        sin2u = (cosu + cosu) * sinu
        cos2u = 1.0 - 2.0 * sinu * sinu
        temp = 1.0 / pl
        temp1 = 0.5 * J2 * temp
        temp2 = temp1 * temp
        mrt = rl * (1.0 - 1.5 * temp2 * betal * g("con41")) + 0.5 * temp1 * g("x1mth2") * cos2u
        su = su - 0.25 * temp2 * g("x7thm1") * sin2u
        xnode = nodem + 1.5 * temp2 * cosim * sin2u
        xinc = g("inclo") + 1.5 * temp2 * cosim * sinim * cos2u
        mvt = rdotl - nm * temp1 * g("x1mth2") * sin2u / XKE
        rvdot = rvdotl + nm * temp1 * (g("x1mth2") * cos2u + 1.5 * g("con41")) / XKE
        sinsu, cossu = np.sin(su), np.cos(su)
        snod, cnod = np.sin(xnode), np.cos(xnode)
        sini, cosi = np.sin(xinc), np.cos(xinc)
        xmx = -snod * cosi
        xmy = cnod * cosi
        ux = xmx * sinsu + cnod * cossu
        uy = xmy * sinsu + snod * cossu
        uz = sini * sinsu
        vx = xmx * cossu - cnod * sinsu
        vy = xmy * cossu - snod * sinsu
        vz = sini * cossu
        mr = mrt * RE
        r = np.stack([mr * ux, mr * uy, mr * uz], axis=-1)
        v = np.stack([(mvt * ux + rvdot * vx) * VKMPERSEC,
                      (mvt * uy + rvdot * vy) * VKMPERSEC,
                      (mvt * uz + rvdot * vz) * VKMPERSEC], axis=-1)
        err |= np.where(mrt < 1.0, ERR_DECAYED, 0).astype(np.int32)
    if return_mean:
        return r, v, err, mean
    return r, v, err


# ---------------------------------------------------------------------------
# epochs: dsgp4.util.from_datetime_to_mjd(tle._epoch) (util.py:570-575)
# ---------------------------------------------------------------------------
def jday(year, mon, day, hr, minute, sec):
    """Vallado ``jday`` -> (jd, fraction)."""
    jd = (367.0 * year - math.floor((7 * (year + math.floor((mon + 9) / 12.0))) * 0.25)
          + math.floor(275 * mon / 9.0) + day + 1721013.5)
    fr = (sec + minute * 60.0 + hr * 3600.0) / 86400.0
    return jd, fr


def datetime_to_mjd(dt):
    jd, fr = jday(dt.year, dt.month, dt.day, dt.hour, dt.minute, dt.second + dt.microsecond / 1e6)
    return jd + fr - 2400000.5


def days2mdhms(year, days):
    """Vallado ``days2mdhms``: day-of-year (1-based, fractional) -> m,d,h,m,s."""
    lmonth = [31, 29 if year % 4 == 0 else 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]
    dayofyr = int(math.floor(days))
    i = 1
    inttemp = 0
    while dayofyr > inttemp + lmonth[i - 1] and i < 12:
        inttemp += lmonth[i - 1]
        i += 1
    mon = i
    day = dayofyr - inttemp
    temp = (days - dayofyr) * 24.0
    hr = int(math.floor(temp))
    temp = (temp - hr) * 60.0
    minute = int(math.floor(temp))
    sec = (temp - minute) * 60.0
    return mon, day, hr, minute, sec


def tle_epoch_datetime(epoch_year, epoch_days):
    """TLE epoch -> microsecond-TRUNCATED datetime, the ``tle._epoch`` the
    reference subtracts at util.py:570,575."""
    mon, day, hr, minute, sec = days2mdhms(epoch_year, epoch_days)
    sec_whole = int(math.floor(sec))
    micro = int((sec - sec_whole) * 1e6)
    if micro > 999999:
        micro = 999999
    return datetime.datetime(epoch_year, mon, day, hr, minute, sec_whole, micro)


def parse_tle_lines(line1, line2):
    """Minimal TLE text -> SGP4 element dict (units: rad, rad/min)."""
    def _exp_field(s):
        s = s.strip()
        if not s or set(s) <= set("+-0 "):
            return 0.0
        mant, ex = s[:-2], s[-2:]
        sign = -1.0 if mant.startswith("-") else 1.0
        mant = mant.lstrip("+-").strip()
        return sign * float("0." + mant) * 10.0 ** int(ex)
    yy = int(line1[18:20])
    year = 2000 + yy if yy < 57 else 1900 + yy
    epoch_days = float(line1[20:32])
    deg = math.pi / 180.0
    rev_day = float(line2[52:63])
    d = dict(
        epoch_year=year, epoch_days=epoch_days,
        bstar=_exp_field(line1[53:61]),
        inclo=float(line2[8:16]) * deg,
        nodeo=float(line2[17:25]) * deg,
        ecco=float("0." + line2[26:33].strip()),
        argpo=float(line2[34:42]) * deg,
        mo=float(line2[43:51]) * deg,
        no_kozai=rev_day * TWOPI / 86400.0 * 60.0,
        mean_motion=rev_day * TWOPI / 86400.0,
    )
    d["epoch_dt"] = tle_epoch_datetime(year, epoch_days)
    d["epoch_mjd"] = datetime_to_mjd(d["epoch_dt"])
    return d
