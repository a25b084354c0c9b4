"""Conjunction Data Message record and its KVN wire format.

The record behind ``generate_cdm`` / ``events.to_cdms`` with the accessors of
/root/reference/kessler/cdm.py (``set_state`` recomputes the relative RTN state and the miss
distance, cdm.py:249-303) and the CCSDS 508.0-B-1 key-value notation the reference reads and
writes (cdm.py:129-183, 391-418): same keys in the same order, ``KEY`` padded to 37 columns,
``KEY = value`` lines, obligatory keys always present, numbers printed with ``str`` unless that
gives an exponent (then ``%.3E``) -- so a file written here loads in Kessler and vice versa.
States and the 6x6 covariances live in arrays (the batched ground segment fills them from
struct-of-arrays results); everything else is a plain ordered mapping.  pandas export is the
one-row frame of ``to_dict()``; the Kelvins loaders and plotting stay out of scope
(SURVEY.md Sec. 2 rows 9-10)."""
import copy
import datetime

import numpy as np

STATE_KEYS = ("X", "Y", "Z", "X_DOT", "Y_DOT", "Z_DOT")
_RTN = ("R", "T", "N", "RDOT", "TDOT", "NDOT")
# lower-triangle key of covariance element (i, j), i >= j: CR_R, CT_R, CT_T, CN_R, ...
COV_KEYS = {(i, j): "C{}_{}".format(_RTN[i], _RTN[j]) for i in range(6) for j in range(i + 1)}
_COV_INDEX = {name: ij for ij, name in COV_KEYS.items()}
# the optional rows of the 9x9 covariance (drag, solar radiation pressure, thrust): carried, never computed here
COV_EXTRA_KEYS = tuple("C{}_{}".format(row, col)
                       for row, cols in (("DRG", _RTN + ("DRG",)), ("SRP", _RTN + ("DRG", "SRP")),
                                         ("THR", _RTN + ("DRG", "SRP", "THR")))
                       for col in cols)
HEADER_KEYS = ("CCSDS_CDM_VERS", "CREATION_DATE", "ORIGINATOR", "MESSAGE_FOR", "MESSAGE_ID")
RELATIVE_KEYS = ("TCA", "MISS_DISTANCE", "RELATIVE_SPEED", "RELATIVE_POSITION_R", "RELATIVE_POSITION_T",
                 "RELATIVE_POSITION_N", "RELATIVE_VELOCITY_R", "RELATIVE_VELOCITY_T", "RELATIVE_VELOCITY_N",
                 "START_SCREEN_PERIOD", "STOP_SCREEN_PERIOD", "SCREEN_VOLUME_FRAME", "SCREEN_VOLUME_SHAPE",
                 "SCREEN_VOLUME_X", "SCREEN_VOLUME_Y", "SCREEN_VOLUME_Z", "SCREEN_ENTRY_TIME", "SCREEN_EXIT_TIME",
                 "COLLISION_PROBABILITY", "COLLISION_PROBABILITY_METHOD")
METADATA_KEYS = ("OBJECT", "OBJECT_DESIGNATOR", "CATALOG_NAME", "OBJECT_NAME", "INTERNATIONAL_DESIGNATOR", "OBJECT_TYPE",
                 "OPERATOR_CONTACT_POSITION", "OPERATOR_ORGANIZATION", "OPERATOR_PHONE", "OPERATOR_EMAIL",
                 "EPHEMERIS_NAME", "COVARIANCE_METHOD", "MANEUVERABLE", "ORBIT_CENTER", "REF_FRAME", "GRAVITY_MODEL",
                 "ATMOSPHERIC_MODEL", "N_BODY_PERTURBATIONS", "SOLAR_RAD_PRESSURE", "EARTH_TIDES", "INTRACK_THRUST")
OD_KEYS = ("TIME_LASTOB_START", "TIME_LASTOB_END", "RECOMMENDED_OD_SPAN", "ACTUAL_OD_SPAN", "OBS_AVAILABLE", "OBS_USED",
           "TRACKS_AVAILABLE", "TRACKS_USED", "RESIDUALS_ACCEPTED", "WEIGHTED_RMS", "AREA_PC", "AREA_DRG", "AREA_SRP", "MASS",
           "CD_AREA_OVER_MASS", "CR_AREA_OVER_MASS", "THRUST_ACCELERATION", "SEDR")
OBLIGATORY = {"header": ("CCSDS_CDM_VERS", "CREATION_DATE", "ORIGINATOR", "MESSAGE_ID"),
              "relative metadata": ("TCA", "MISS_DISTANCE"),
              "metadata": ("OBJECT", "OBJECT_DESIGNATOR", "CATALOG_NAME", "OBJECT_NAME", "INTERNATIONAL_DESIGNATOR",
                           "EPHEMERIS_NAME", "COVARIANCE_METHOD", "MANEUVERABLE", "REF_FRAME"),
              "data (od)": (),
              "data (state)": STATE_KEYS,
              "data (covariance)": tuple(COV_KEYS.values())}
_KEY_COLUMNS = 37


def _unset(v):
    return v is None or (isinstance(v, float) and v != v)


def _is_number(text):
    try:
        float(text)
        return True
    except ValueError:
        return False


def _format_value(v):
    if isinstance(v, (float, int, np.floating, np.integer)) and not isinstance(v, bool):
        text = "{}".format(float(v) if isinstance(v, np.floating) else v)
        return "{:.3E}".format(v) if "e" in text else text
    return str(v)


# header keys that carry a CCSDS time string (cdm.py:53,193-203: day-of-year timestamps are converted, anything that is
# not yyyy-mm-ddTHH:MM:SS.f afterwards is refused)
_DATE_KEYS = ("CREATION_DATE",)


def _ccsds_date(key, value):
    text = str(value)
    if text.count("T") != 1 or text.find("T") not in (8, 10):
        raise RuntimeError("*** Error -- Invalid CCSDS time string: {}\nDate format not one of yyyy-mm-dd or yyyy-DDD.".format(text))
    if text.find("T") == 8:      # yyyy-DDDTHH:MM:SS[.f][Z]
        day = datetime.datetime(int(text[0:4]), 1, 1) + datetime.timedelta(int(text[5:8]) - 1)
        text = "{:04d}-{:02d}-{:02d}T{}".format(day.year, day.month, day.day, text[9:].rstrip("Z"))
    try:
        datetime.datetime.strptime(text, "%Y-%m-%dT%H:%M:%S.%f")
    except Exception as e:
        raise RuntimeError("{} ({}) is not in the expected format.\n{}".format(key, text, e))
    return text


class ConjunctionDataMessage:
    def __init__(self, file_name=None, set_defaults=True):
        self._header = dict.fromkeys(HEADER_KEYS)
        self._relative = dict.fromkeys(RELATIVE_KEYS)
        self._meta = [dict.fromkeys(METADATA_KEYS), dict.fromkeys(METADATA_KEYS)]
        self._od = [dict.fromkeys(OD_KEYS), dict.fromkeys(OD_KEYS)]
        self._cov_extra = [dict.fromkeys(COV_EXTRA_KEYS), dict.fromkeys(COV_EXTRA_KEYS)]
        self._state = np.full((2, 2, 3), np.nan)       # [object, pos|vel, xyz]  km, km/s (ITRF)
        self._cov = np.full((2, 6, 6), np.nan)         # [object] RTN, m^2 ...
        if set_defaults:
            self._header["CCSDS_CDM_VERS"] = "1.0"
            self._header["CREATION_DATE"] = datetime.datetime.now(datetime.timezone.utc).replace(tzinfo=None).isoformat()
            self._meta[0]["OBJECT"], self._meta[1]["OBJECT"] = "OBJECT1", "OBJECT2"
        if file_name:
            self.copy_from(ConjunctionDataMessage.load(file_name))

    def copy(self):
        return copy.deepcopy(self)

    def copy_from(self, other):
        self.__dict__.update(copy.deepcopy(other.__dict__))

    # -- setters / getters with the reference's names --------------------------------------
    def set_header(self, key, value):
        if key not in HEADER_KEYS:
            raise ValueError("Invalid key ({}) for header".format(key))
        if key in _DATE_KEYS:
            value = _ccsds_date(key, value)
        self._header[key] = value

    def set_relative_metadata(self, key, value):
        if key not in RELATIVE_KEYS:
            raise ValueError("Invalid key ({}) for relative metadata".format(key))
        self._relative[key] = value

    def get_relative_metadata(self, key):
        if key not in RELATIVE_KEYS:
            raise ValueError("Invalid key ({}) for relative metadata".format(key))
        return self._relative[key]

    @staticmethod
    def _check_id(object_id):
        if object_id != 0 and object_id != 1:
            raise ValueError("Expecting object_id to be 0 or 1")

    def set_object(self, object_id, key, value):
        self._check_id(object_id)
        if key in METADATA_KEYS:
            self._meta[object_id][key] = value
        elif key in OD_KEYS:
            self._od[object_id][key] = value
        elif key in STATE_KEYS:
            k = STATE_KEYS.index(key)
            self._state[object_id, k // 3, k % 3] = np.nan if value is None else value   # (only set_state refreshes the relative state)
        elif key in _COV_INDEX:
            i, j = _COV_INDEX[key]
            self._cov[object_id, i, j] = self._cov[object_id, j, i] = np.nan if value is None else value
        elif key in COV_EXTRA_KEYS:
            self._cov_extra[object_id][key] = value
        else:
            raise ValueError("Invalid key ({}) for object data".format(key))

    def get_object(self, object_id, key):
        self._check_id(object_id)
        if key in METADATA_KEYS:
            return self._meta[object_id][key]
        if key in OD_KEYS:
            return self._od[object_id][key]
        if key in STATE_KEYS:
            k = STATE_KEYS.index(key)
            return self._state[object_id, k // 3, k % 3]
        if key in _COV_INDEX:
            i, j = _COV_INDEX[key]
            return self._cov[object_id, i, j]
        if key in COV_EXTRA_KEYS:
            return self._cov_extra[object_id][key]
        raise ValueError("Invalid key ({}) for object data".format(key))

    def _refresh_relative(self):
        s1, s2 = self._state[0] * 1e3, self._state[1] * 1e3
        r, v = s1[0], s1[1]
        with np.errstate(all="ignore"):
            u = r / np.linalg.norm(r)
            w = np.cross(r, v)
            w = w / np.linalg.norm(w)
            rot = np.vstack((u, np.cross(w, u), w))
            rel_p, rel_v = rot @ (s2[0] - s1[0]), rot @ (s2[1] - s1[1])
            for axis, name in enumerate("RTN"):
                self._relative["RELATIVE_POSITION_" + name] = rel_p[axis]
                self._relative["RELATIVE_VELOCITY_" + name] = rel_v[axis]
            self._relative["RELATIVE_SPEED"] = np.linalg.norm(rel_v)
            self._relative["MISS_DISTANCE"] = np.linalg.norm(s1[0] - s2[0])

    def set_state(self, object_id, state):
        """state [2,3] km, km/s; refreshes the relative RTN state and the miss distance."""
        self._check_id(object_id)
        self._state[object_id] = np.asarray(state, dtype=np.float64)
        self._refresh_relative()

    def get_state(self, object_id):
        self._check_id(object_id)
        return self._state[object_id].copy()

    def get_state_relative(self):
        g = self._relative
        return np.array([[g["RELATIVE_POSITION_" + a] for a in "RTN"], [g["RELATIVE_VELOCITY_" + a] for a in "RTN"]], dtype=np.float64)

    def set_covariance(self, object_id, covariance_matrix):
        self._check_id(object_id)
        c = np.asarray(covariance_matrix, dtype=np.float64)
        low = np.tril(c)
        self._cov[object_id] = low + low.T - np.diag(np.diag(c))     # the record keeps the lower triangle

    def get_covariance(self, object_id):
        self._check_id(object_id)
        return self._cov[object_id].copy()

    # -- dict-style access by CDM key (cdm['TCA'], cdm['OBJECT1_X'], ...) ----------------------
    def __getitem__(self, key):
        if key in HEADER_KEYS:
            return self._header[key]
        if key in RELATIVE_KEYS:
            return self._relative[key]
        for oid, prefix in enumerate(("OBJECT1_", "OBJECT2_")):
            if key.startswith(prefix):
                return self.get_object(oid, key[len(prefix):])
        raise ValueError("Invalid key: {}".format(key))

    def __setitem__(self, key, value):
        if key in HEADER_KEYS:
            self.set_header(key, value)
        elif key in RELATIVE_KEYS:
            self.set_relative_metadata(key, value)
        elif key.startswith("OBJECT1_"):
            self.set_object(0, key[len("OBJECT1_"):], value)
        elif key.startswith("OBJECT2_"):
            self.set_object(1, key[len("OBJECT2_"):], value)
        else:
            raise ValueError("Invalid key: {}".format(key))

    # -- the record as ordered sections: (part name, obligatory keys, {key: value}) -------------
    def _object_sections(self, oid):
        state = {k: (None if _unset(self._state[oid, i // 3, i % 3]) else float(self._state[oid, i // 3, i % 3]))
                 for i, k in enumerate(STATE_KEYS)}
        cov = {name: (None if _unset(self._cov[oid, i, j]) else float(self._cov[oid, i, j])) for (i, j), name in COV_KEYS.items()}
        cov.update(self._cov_extra[oid])
        return (("metadata", self._meta[oid]), ("data (od)", self._od[oid]), ("data (state)", state), ("data (covariance)", cov))

    def _sections(self):
        yield "header", "header", self._header
        yield "relative metadata", "relative metadata", self._relative
        for oid in (0, 1):
            for part, values in self._object_sections(oid):
                yield "Object{} {}".format(oid + 1, part), part, values

    def to_dict(self):
        d = dict(self._header)
        d.update({k: (None if _unset(v) else v) for k, v in self._relative.items()})
        for oid, prefix in enumerate(("OBJECT1_", "OBJECT2_")):
            for _, values in self._object_sections(oid):
                for k, v in values.items():
                    d[prefix + k] = v
        return d

    def to_dataframe(self):
        import pandas as pd
        return pd.DataFrame(self.to_dict(), index=[0])

    def validate(self):
        """Prints the obligatory keys that have no value (cdm.py:375-389) and returns them."""
        missing = []
        for label, part, values in self._sections():
            for key in OBLIGATORY[part]:
                if _unset(values[key]):
                    print("Missing obligatory value in {}: {}".format(label, key))
                    missing.append((label, key))
        return missing

    def kvn(self, show_all=False):
        """The message in key-value notation: set keys, plus the obligatory ones (all of them with show_all) left empty."""
        lines = []
        for _, part, values in self._sections():
            for k, v in values.items():
                if _unset(v):
                    if show_all or k in OBLIGATORY[part]:
                        lines.append("{} =\n".format(k.ljust(_KEY_COLUMNS)))
                else:
                    lines.append("{} = {}\n".format(k.ljust(_KEY_COLUMNS), _format_value(v)))
        return "".join(lines)

    def save(self, file_name):
        with open(file_name, "w") as f:
            f.write(self.kvn())

    @staticmethod
    def load(file_name):
        """Reads a KVN file (cdm.py:129-178): COMMENT and empty lines are skipped, numeric values become floats, keys that
        follow OBJECT = OBJECT1 / OBJECT2 belong to that object, unknown keys are ignored."""
        cdm = ConjunctionDataMessage(set_defaults=False)
        target = None          # None: header / relative metadata; 0 / 1: the object being read
        with open(file_name) as f:
            for raw in f:
                line = raw.replace("\ufeff", "").strip()
                if not line or line.startswith("COMMENT") or "=" not in line:
                    continue
                key, value = (part.strip() for part in line.split("=", 1))
                if value == "":
                    continue
                value = float(value) if _is_number(value) else value
                if key == "OBJECT" and value in ("OBJECT1", "OBJECT2"):
                    target = 0 if value == "OBJECT1" else 1
                    cdm._meta[target]["OBJECT"] = value
                elif target is None:
                    if key in HEADER_KEYS:
                        cdm.set_header(key, value)
                    elif key in RELATIVE_KEYS:
                        cdm.set_relative_metadata(key, value)
                else:
                    try:
                        cdm.set_object(target, key, value)
                    except ValueError:
                        continue
        return cdm

    def __hash__(self):
        return hash(self.kvn(show_all=True))

    def __eq__(self, other):
        return isinstance(other, ConjunctionDataMessage) and self.kvn(show_all=True) == other.kvn(show_all=True)

    def __repr__(self):
        return self.kvn()


CDM = ConjunctionDataMessage


class Event:
    """A time series of CDMs of one conjunction (util.trace_to_event); list-like."""

    def __init__(self, cdms=None):
        self._cdms = list(cdms or [])

    def __len__(self):
        return len(self._cdms)

    def __getitem__(self, i):
        return self._cdms[i]

    def __iter__(self):
        return iter(self._cdms)

    def to_dataframe(self):
        """One row per CDM, columns = CDM keys (event.py:64-72)."""
        import pandas as pd
        if not self._cdms:
            return pd.DataFrame()
        return pd.concat([c.to_dataframe() for c in self._cdms], ignore_index=True)

    def save(self, directory, prefix="event_0", extension=".kvn"):
        """Writes one KVN file per CDM, <directory>/<prefix>_<k><extension> -- the naming the reference's
        EventDataset(cdms_dir=directory) groups into events (event.py:168-189: `(.*)_([0-9]+.kvn)`); returns the paths."""
        import os
        os.makedirs(directory, exist_ok=True)
        paths = []
        for k, c in enumerate(self._cdms):
            paths.append(os.path.join(directory, "{}_{}{}".format(prefix, k, extension)))
            c.save(paths[-1])
        return paths
