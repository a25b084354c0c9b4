"""Stand-in for ``dsgp4.tle.TLE`` -- the part of dsgp4's object model that Kessler's hot
path touches (``TLE(list|dict)``, ``.copy()``, ``.update(dict)``,
``.apogee_alt()``, ``.perigee_alt()``, the element attributes, ``line1/line2`` and ``load``).

dsgp4 (esa/dSGP4) is an un-vendored, un-pinned dependency of the reference
(/root/reference/setup.py:35) and is not importable here, so this is a from-scratch
implementation of the TLE text format with dsgp4's units and attribute names.  Pinned by:
* docs/notebooks/trace.pickle: every parsed field of two TLEs bit-exactly, including the
  text round trip ``update({'mean_anomaly': m})`` performs (sample 0.4932476473698829 ->
  ``_mo`` 0.49324749990611744 = 28.2610 deg), tests/test_tle.py;
* /root/reference/tests/test_model.py:19-39: ``copy()`` / ``update()`` reproduce line1/line2.

Units (dsgp4): ``mean_motion`` rad/s, ``mean_motion_first_derivative`` rad/s^2 (the full
n-dot: 2 x the TLE field), ``mean_motion_second_derivative`` rad/s^3 (6 x the field), angles
rad, ``b_star`` 1/earth-radii; ``_no_kozai`` rad/min.
"""
import copy as _copy
import datetime
import math

import numpy as np

MU_EARTH = 398600.5e9        # m^3/s^2, WGS-84 as dsgp4.util.get_gravity_constants('wgs-84') * 1e9
R_EARTH = 6378.137e3         # m
_TWO_PI = 2.0 * math.pi
_DEG = math.pi / 180.0


# ---- epoch helpers (Vallado jday / days2mdhms, as dsgp4.util) ---------------------------
def jday(year, mon, day, hr, minute, sec):
    jd = (367.0 * year - math.floor((7 * (year + math.floor((mon + 9) / 12.0))) * 0.25)
          + math.floor(275 * mon / 9.0) + day + 1721013.5)
    fr = (sec + minute * 60.0 + hr * 3600.0) / 86400.0
    return jd, fr


def from_datetime_to_mjd(dt):
    jd, fr = jday(dt.year, dt.month, dt.day, dt.hour, dt.minute, dt.second + dt.microsecond / 1e6)
    return jd + fr - 2400000.5


def from_datetime_to_jd(dt):
    jd, fr = jday(dt.year, dt.month, dt.day, dt.hour, dt.minute, dt.second + dt.microsecond / 1e6)
    return jd + fr


def days2mdhms(year, days):
    lmonth = [31, 29 if year % 4 == 0 else 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]
    dayofyr = int(math.floor(days))
    i, inttemp = 1, 0
    while dayofyr > inttemp + lmonth[i - 1] and i < 12:
        inttemp += lmonth[i - 1]
        i += 1
    temp = (days - dayofyr) * 24.0
    hr = int(math.floor(temp))
    temp = (temp - hr) * 60.0
    minute = int(math.floor(temp))
    sec = (temp - minute) * 60.0
    return i, dayofyr - inttemp, hr, minute, sec


def invjday(jd):
    """Julian date -> (year, mon, day, hr, minute, sec), Vallado."""
    temp = jd - 2415019.5
    tu = temp / 365.25
    year = 1900 + int(math.floor(tu))
    leapyrs = int(math.floor((year - 1901) * 0.25))
    days = temp - ((year - 1900) * 365.0 + leapyrs) + 0.00000000001
    if days < 1.0:
        year -= 1
        leapyrs = int(math.floor((year - 1901) * 0.25))
        days = temp - ((year - 1900) * 365.0 + leapyrs)
    mon, day, hr, minute, sec = days2mdhms(year, days)
    return year, mon, day, hr, minute, sec - 0.00000086400


def from_jd_to_datetime(jd):
    year, mon, day, hr, minute, sec = invjday(jd)
    sec = max(sec, 0.0)
    whole = int(math.floor(sec))
    micro = int((sec - whole) * 1e6)
    return datetime.datetime(year, mon, day, hr, minute, min(whole, 59), min(micro, 999999))


def from_mjd_to_datetime(mjd):
    return from_jd_to_datetime(mjd + 2400000.5)


def from_mjd_to_epoch_days_after_1_jan(mjd):
    d = from_mjd_to_datetime(mjd)
    dd = d - datetime.datetime(d.year - 1, 12, 31)
    return dd.days + (dd.seconds + dd.microseconds / 1e6) / 86400.0


def epoch_to_datetime(epoch_year, epoch_days):
    """TLE epoch -> datetime truncated to the microsecond (the ``_epoch`` Kessler subtracts at
    util.py:570,575).  Its MJD goes through jd + fr at magnitude 2.46e6, i.e. a 40 us grid,
    so a +-1 us difference in this truncation is not observable downstream."""
    mon, day, hr, minute, sec = days2mdhms(epoch_year, epoch_days)
    whole = int(math.floor(sec))
    micro = min(int((sec - whole) * 1e6), 999999)
    return datetime.datetime(epoch_year, mon, day, hr, minute, whole, micro)


# ---- TLE text fields -----------------------------------------------------------------------
def _checksum(line68):
    s = 0
    for ch in line68:
        if ch.isdigit():
            s += int(ch)
        elif ch == "-":
            s += 1
    return s % 10


def _parse_exp_field(s):
    """' 41303-3' -> 0.41303e-3 ; ' 00000-0' -> 0."""
    s = s.strip()
    if not s:
        return 0.0
    sign = -1.0 if s[0] == "-" else 1.0
    s = s.lstrip("+-")
    mant, ex = s[:-2], s[-2:]
    if not mant.strip("0"):
        return 0.0
    return sign * float("0." + mant) * 10.0 ** int(ex)


def _format_exp_field(x):
    """0.00041303 -> ' 41303-3' (8 chars, implied leading decimal point)."""
    x = float(x)
    if x == 0.0 or not math.isfinite(x):
        return " 00000-0"
    sign = "-" if x < 0 else " "
    ax = abs(x)
    ex = int(math.floor(math.log10(ax))) + 1
    mant = int(round(ax / 10.0 ** ex * 1e5))
    if mant >= 100000:
        mant //= 10
        ex += 1
    if mant == 0 or ex < -9:       # below the field's range (one exponent digit): reads back as zero
        return " 00000-0"
    if ex > 9:                      # above it: saturate
        mant, ex = 99999, 9
    return "{}{:05d}{}{:d}".format(sign, mant, "-" if ex <= 0 else "+", abs(ex))


def _format_ndot_field(x):
    """first derivative field: ' .00010731' / '-.00010731' (10 chars)."""
    s = "{:.8f}".format(abs(float(x)))
    s = s[1:] if s.startswith("0") else s
    return ("-" if x < 0 else " ") + s[:9]


class TLE:
    """Two-line element set with dsgp4's attribute surface."""

    def __init__(self, data):
        if isinstance(data, TLE):
            data = [l for l in (data.line0, data.line1, data.line2) if l]
        if isinstance(data, dict):
            d = dict(data)
            self._from_fields(d, name=d.get("name"))
        else:
            lines = [str(l).rstrip("\n") for l in data]
            name = None
            if len(lines) == 3:
                name = lines[0][2:].strip() if lines[0].startswith("0 ") else lines[0].strip()
                lines = lines[1:]
            if len(lines) != 2 or not lines[0].startswith("1 ") or not lines[1].startswith("2 "):
                raise ValueError("TLE: expected [line1, line2] or [line0, line1, line2]")
            self._from_lines(lines[0], lines[1], name)

    # -- construction -------------------------------------------------------------------
    def _from_lines(self, line1, line2, name=None):
        yy = int(line1[18:20])
        d = dict(
            satellite_catalog_number=int(line1[2:7]),
            classification=line1[7],
            international_designator=line1[9:17].strip(),
            epoch_year=2000 + yy if yy < 57 else 1900 + yy,
            epoch_days=float(line1[20:32]),
            mean_motion_first_derivative=float(line1[33:43].replace(" ", "")) * 2.0 * _TWO_PI / 86400.0 ** 2,
            mean_motion_second_derivative=_parse_exp_field(line1[44:52]) * 6.0 * _TWO_PI / 86400.0 ** 3,
            b_star=_parse_exp_field(line1[53:61]),
            ephemeris_type=int(line1[62]) if line1[62].strip() else 0,
            element_number=int(line1[64:68]),
            inclination=float(line2[8:16]) * _DEG,
            raan=float(line2[17:25]) * _DEG,
            eccentricity=float("0." + line2[26:33].replace(" ", "0")),
            argument_of_perigee=float(line2[34:42]) * _DEG,
            mean_anomaly=float(line2[43:51]) * _DEG,
            mean_motion=float(line2[52:63]) * _TWO_PI / 86400.0,
            revolution_number_at_epoch=int(line2[63:68]),
        )
        self._set(d, line1, line2, name)

    def _from_fields(self, d, name=None):
        """dict -> canonical text -> parse: the TLE text quantises every element
        (SURVEY.md Sec. 8a row 1), so go through it exactly as ``update`` does."""
        f = lambda k: float(d[k])
        yr = int(d["epoch_year"])
        l1 = "1 {:05d}{} {:<8s} {:02d}{:012.8f} {} {} {} {:1d} {:4d}".format(
            int(d["satellite_catalog_number"]), str(d.get("classification", "U"))[:1],
            str(d["international_designator"])[:8], yr % 100, f("epoch_days"),
            _format_ndot_field(f("mean_motion_first_derivative") * 86400.0 ** 2 / (2.0 * _TWO_PI)),
            _format_exp_field(f("mean_motion_second_derivative") * 86400.0 ** 3 / (6.0 * _TWO_PI)),
            _format_exp_field(f("b_star")), int(d.get("ephemeris_type", 0)), int(d.get("element_number", 999)) % 10000)
        ecc = "{:.7f}".format(f("eccentricity"))
        ecc = "9999999" if ecc.startswith("1") else ecc[2:9]
        l2 = "2 {:05d} {:8.4f} {:8.4f} {} {:8.4f} {:8.4f} {:11.8f}{:5d}".format(
            int(d["satellite_catalog_number"]), f("inclination") / _DEG, f("raan") / _DEG, ecc,
            f("argument_of_perigee") / _DEG, f("mean_anomaly") / _DEG, f("mean_motion") * 86400.0 / _TWO_PI,
            int(d.get("revolution_number_at_epoch", 0)) % 100000)
        l1 = l1[:68].ljust(68)
        l2 = l2[:68].ljust(68)
        self._from_lines(l1 + str(_checksum(l1)), l2 + str(_checksum(l2)), name)
        # dsgp4 keeps the caller's values in the public fields and derives only the SGP4 inputs
        # (_mo, _no_kozai, ...) from the text: after update({'mean_anomaly': 0.4932476473698829}) the
        # pickled golden TLE has mean_anomaly = that sample but _mo = 0.49324749990611744 (28.2610 deg)
        for k, v in d.items():
            if k in self._data and k not in ("epoch_year", "epoch_days"):
                self._data[k] = v
                setattr(self, k, v)
        if float(d["mean_motion"]) > 0:
            self.semi_major_axis = (MU_EARTH / float(d["mean_motion"]) ** 2) ** (1.0 / 3.0)

    def _set(self, d, line1, line2, name):
        self._data = d
        self._lines = [line1, line2]
        self.line1, self.line2 = line1, line2
        self.line0 = ("0 " + name) if name else None
        if name:
            self.name = name
        for k, v in d.items():
            setattr(self, k, v)
        self._epoch = epoch_to_datetime(d["epoch_year"], d["epoch_days"])
        self.date_mjd = from_datetime_to_mjd(self._epoch)
        # SGP4 inputs (rad, rad/min), as dsgp4 names them
        self._no_kozai = d["mean_motion"] * 60.0
        self._ecco = d["eccentricity"]
        self._inclo = d["inclination"]
        self._nodeo = d["raan"]
        self._argpo = d["argument_of_perigee"]
        self._mo = d["mean_anomaly"]
        self._bstar = d["b_star"]
        # Vallado twoline2rv scaling of the raw text fields (not used by near-earth SGP4)
        xpdotp = 1440.0 / _TWO_PI
        self._ndot = float(line1[33:43].replace(" ", "")) / (xpdotp * 1440.0)
        self._nddot = _parse_exp_field(line1[44:52]) / (xpdotp * 1440.0 * 1440.0)
        self.semi_major_axis = (MU_EARTH / d["mean_motion"] ** 2) ** (1.0 / 3.0) if d["mean_motion"] > 0 else float("nan")
        self._initialized = False

    # -- dsgp4 surface ------------------------------------------------------------------
    def copy(self):
        """Same object state, verbatim: the public fields (which keep the caller's raw values after a dict
        construction / update) AND the text-quantised SGP4 inputs (_no_kozai, _ecco, ... -- what propagates).
        Re-deriving the latter from the public fields would drop the quantisation."""
        new = TLE.__new__(TLE)
        for k, v in self.__dict__.items():
            new.__dict__[k] = dict(v) if isinstance(v, dict) else (list(v) if isinstance(v, list) else v)
        new._initialized = False
        return new

    def update(self, fields):
        """Re-encode with the given fields replaced (dsgp4 ``TLE.update``): the result is
        quantised by the TLE text format."""
        d = dict(self._data)
        for k, v in fields.items():
            if k not in d:
                raise KeyError(f"TLE.update: unknown field {k!r}")
            d[k] = float(v) if not isinstance(v, (int, str)) else v
        self._from_fields(d, name=getattr(self, "name", None))
        return self

    def apogee_alt(self, earth_radius=R_EARTH):
        return self.semi_major_axis * (1.0 + self.eccentricity) - earth_radius

    def perigee_alt(self, earth_radius=R_EARTH):
        return self.semi_major_axis * (1.0 - self.eccentricity) - earth_radius

    def elements(self):
        """SGP4 input vector (no_kozai, ecco, inclo, nodeo, argpo, mo, bstar)."""
        return np.array([self._no_kozai, self._ecco, self._inclo, self._nodeo, self._argpo, self._mo, self._bstar],
                        dtype=np.float64)

    def __repr__(self):
        return "TLE(\n{}\n{}\n)".format(self.line1, self.line2)

    def __deepcopy__(self, memo):
        return self.copy()


def load_from_lines(lines):
    lines = [l.rstrip("\n") for l in lines if l.strip()]
    out, i = [], 0
    while i < len(lines):
        if lines[i].startswith("1 ") and i + 1 < len(lines) and lines[i + 1].startswith("2 "):
            out.append(TLE(lines[i:i + 2]))
            i += 2
        elif i + 2 < len(lines) and lines[i + 1].startswith("1 ") and lines[i + 2].startswith("2 "):
            out.append(TLE(lines[i:i + 3]))
            i += 3
        else:
            raise ValueError(f"TLE file: cannot parse at line {i + 1}: {lines[i]!r}")
    return out


def load(file_name):
    """dsgp4.tle.load: list of TLEs from a 2- or 3-line-per-object text file."""
    with open(file_name) as fh:
        return load_from_lines(fh.readlines())


_ = _copy  # (kept for API parity: dsgp4 exposes copy semantics through TLE.copy)
