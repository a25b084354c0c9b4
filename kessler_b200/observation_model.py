"""Observation models of /root/reference/kessler/observation_model.py:14-48: an instrument
returns the true state plus a fixed bias and its fixed diagonal RTN covariance."""
import numpy as np


class ObservationModel:
    def __init__(self, instrument_characteristics=None):
        self._instrument_characteristics = {} if instrument_characteristics is None else instrument_characteristics

    def __repr__(self):
        return "ObservationModel"

    def observe(self, time, spacecraft):
        raise NotImplementedError()


class _BiasedInstrument(ObservationModel):
    _default = None
    _label = "Instrument"

    def __init__(self, instrument_characteristics=None):
        super().__init__(dict(self._default) if instrument_characteristics is None else instrument_characteristics)

    def __repr__(self):
        return "{}({})".format(self._label, self._instrument_characteristics)

    def observe(self, time, spacecraft):
        ic = self._instrument_characteristics
        return spacecraft["state_xyz"] + ic["bias_xyz"], ic["covariance_rtn"]


class GNSS(_BiasedInstrument):
    _default = {"bias_xyz": np.zeros((2, 3)),
                "covariance_rtn": np.array([1e-9, 0.00115849341564346, 0.000059309835843067, 1e-9, 1e-9, 1e-9]) ** 2}
    _label = "GNNS"      # (sic) the reference's repr, observation_model.py:30


class Radar(_BiasedInstrument):
    _default = {"bias_xyz": np.zeros((2, 3)), "covariance_rtn": np.array([20, 500, 1, 0.0001, 0.0001, 0.0001]) ** 2}
    _label = "Radar"
