"""Minimal Conjunction Data Message record: just enough of the container in
/root/reference/kessler/cdm.py for ``generate_cdm`` to put its numbers somewhere and for
users to read them back with the same accessors (``set_state`` recomputes the relative RTN
state and the miss distance, cdm.py:249-303).  The full CCSDS 508.0-B-1 KVN loader / pandas
export / validation of the reference is OUT OF SCOPE."""
import copy
import datetime

import numpy as np

STATE_KEYS = ("X", "Y", "Z", "X_DOT", "Y_DOT", "Z_DOT")
_RTN = ("R", "T", "N", "RDOT", "TDOT", "NDOT")
# lower-triangle key of covariance element (i, j), i >= j: CR_R, CT_R, CT_T, CN_R, ...
COV_KEYS = {(i, j): "C{}_{}".format(_RTN[i], _RTN[j]) for i in range(6) for j in range(i + 1)}
HEADER_KEYS = ("CCSDS_CDM_VERS", "CREATION_DATE", "ORIGINATOR", "MESSAGE_FOR", "MESSAGE_ID")
RELATIVE_KEYS = ("TCA", "MISS_DISTANCE", "RELATIVE_SPEED", "RELATIVE_POSITION_R", "RELATIVE_POSITION_T",
                 "RELATIVE_POSITION_N", "RELATIVE_VELOCITY_R", "RELATIVE_VELOCITY_T", "RELATIVE_VELOCITY_N",
                 "COLLISION_PROBABILITY", "COLLISION_PROBABILITY_METHOD")
METADATA_KEYS = ("OBJECT", "OBJECT_DESIGNATOR", "CATALOG_NAME", "OBJECT_NAME", "INTERNATIONAL_DESIGNATOR", "OBJECT_TYPE",
                 "EPHEMERIS_NAME", "COVARIANCE_METHOD", "MANEUVERABLE", "ORBIT_CENTER", "REF_FRAME")


class ConjunctionDataMessage:
    def __init__(self, set_defaults=True):
        self._header = dict.fromkeys(HEADER_KEYS)
        self._relative = dict.fromkeys(RELATIVE_KEYS)
        self._meta = [dict.fromkeys(METADATA_KEYS), dict.fromkeys(METADATA_KEYS)]
        self._state = np.full((2, 2, 3), np.nan)       # [object, pos|vel, xyz]  km, km/s (ITRF)
        self._cov = np.full((2, 6, 6), np.nan)         # [object] RTN, m^2 ...
        if set_defaults:
            self._header["CCSDS_CDM_VERS"] = "1.0"
            self._header["CREATION_DATE"] = datetime.datetime.now(datetime.timezone.utc).replace(tzinfo=None).isoformat()
            self._meta[0]["OBJECT"], self._meta[1]["OBJECT"] = "OBJECT1", "OBJECT2"

    def copy(self):
        return copy.deepcopy(self)

    # -- setters / getters with the reference's names --------------------------------------
    def set_header(self, key, value):
        if key not in HEADER_KEYS:
            raise ValueError("Invalid key ({}) for header".format(key))
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
        elif key in STATE_KEYS:
            k = STATE_KEYS.index(key)
            self._state[object_id, k // 3, k % 3] = np.nan if value is None else value
        elif key in COV_KEYS.values():
            i, j = next(ij for ij, name in COV_KEYS.items() if name == key)
            self._cov[object_id, i, j] = self._cov[object_id, j, i] = value
        else:
            raise ValueError("Invalid key ({}) for object data".format(key))

    def get_object(self, object_id, key):
        self._check_id(object_id)
        if key in METADATA_KEYS:
            return self._meta[object_id][key]
        if key in STATE_KEYS:
            k = STATE_KEYS.index(key)
            return self._state[object_id, k // 3, k % 3]
        if key in COV_KEYS.values():
            i, j = next(ij for ij, name in COV_KEYS.items() if name == key)
            return self._cov[object_id, i, j]
        raise ValueError("Invalid key ({}) for object data".format(key))

    def set_state(self, object_id, state):
        """state [2,3] km, km/s; refreshes the relative RTN state and the miss distance."""
        self._check_id(object_id)
        self._state[object_id] = np.asarray(state, dtype=np.float64)
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

    def to_dict(self):
        d = dict(self._header)
        d.update(self._relative)
        for oid, prefix in enumerate(("OBJECT1_", "OBJECT2_")):
            for k, v in self._meta[oid].items():
                d[prefix + k] = v
            for k in STATE_KEYS:
                d[prefix + k] = self.get_object(oid, k)
            for (i, j), name in COV_KEYS.items():
                d[prefix + name] = self._cov[oid, i, j]
        return d

    def kvn(self):
        return "\n".join("{} = {}".format(k, v) for k, v in self.to_dict().items() if v is not None)

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
