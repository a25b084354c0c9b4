"""The three things Conjunction.forward() needs from a probabilistic-programming runtime.

The reference records its generative model with Pyro (``pyro.sample``, ``pyro.deterministic``,
``pyro.poutine.trace(fn).get_trace()``; /root/reference/kessler/model.py:15-16,750).  When
pyro is importable it is used unchanged.  It is not installed in this image, so this module
provides a minimal recorder with the same call signatures and the same trace layout
(``trace.nodes[name]['value']`` / ``['infer']``), which is all Kessler and its users read
(SURVEY.md Appendix B).  It performs no inference.
"""
from collections import OrderedDict

import torch

try:
    import pyro as _pyro
    import pyro.distributions as dist
    import pyro.poutine as poutine
    sample = _pyro.sample
    deterministic = _pyro.deterministic
    HAVE_PYRO = True
except Exception:
    HAVE_PYRO = False
    import torch.distributions as _td

    class _Delta(_td.Distribution):
        arg_constraints = {}
        has_rsample = False

        def __init__(self, v, log_density=0.0):
            self.v = torch.as_tensor(v)
            super().__init__(batch_shape=self.v.shape, validate_args=False)

        def sample(self, sample_shape=torch.Size()):
            return self.v.expand(torch.Size(sample_shape) + self.v.shape)

        def log_prob(self, value):
            return torch.where(torch.as_tensor(value) == self.v, torch.zeros(()), torch.full((), -float("inf")))

    class dist:  # namespace: the distributions the model uses
        MixtureSameFamily = _td.MixtureSameFamily
        Categorical = _td.Categorical
        Uniform = _td.Uniform
        Normal = _td.Normal
        Bernoulli = _td.Bernoulli
        Delta = _Delta

    class Trace:
        def __init__(self):
            self.nodes = OrderedDict()

        def add_node(self, name, **kw):
            if name in self.nodes and name not in ("_INPUT", "_RETURN"):
                raise RuntimeError(f"Multiple sample sites named '{name}'")
            self.nodes[name] = dict(name=name, **kw)

        def log_prob_sum(self):
            total = torch.zeros(())
            for n in self.nodes.values():
                if n.get("type") == "sample" and n.get("fn") is not None and not n["infer"].get("_deterministic"):
                    try:
                        total = total + torch.as_tensor(n["fn"].log_prob(torch.as_tensor(n["value"]))).sum()
                    except Exception:
                        pass
            return total

    _STACK = []

    def sample(name, fn, obs=None, infer=None):
        value = obs if obs is not None else fn.sample()
        if _STACK:
            _STACK[-1].add_node(name, type="sample", fn=fn, value=value, is_observed=obs is not None,
                                infer=dict(infer or {}))
        return value

    def deterministic(name, value, event_dim=None):
        value = torch.as_tensor(value) if not torch.is_tensor(value) else value
        if _STACK:
            _STACK[-1].add_node(name, type="sample", fn=None, value=value, is_observed=False,
                                infer={"_deterministic": True})
        return value

    class _TraceHandler:
        def __init__(self, fn):
            self.fn = fn
            self.trace = None

        def get_trace(self, *args, **kwargs):
            tr = Trace()
            tr.add_node("_INPUT", type="args", value=None, args=args, kwargs=kwargs, infer={})
            _STACK.append(tr)
            try:
                ret = self.fn(*args, **kwargs)
            finally:
                _STACK.pop()
            tr.add_node("_RETURN", type="return", value=ret, infer={})
            self.trace = tr
            return tr

        def __call__(self, *args, **kwargs):
            self.get_trace(*args, **kwargs)
            return self.trace.nodes["_RETURN"]["value"]

    class poutine:  # namespace
        trace = _TraceHandler
        Trace = Trace
