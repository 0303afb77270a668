"""Extra distributions of the reference interface (mininf/distributions.py:5-11).

The reference carries a bare-bones ``InverseGamma`` "until pytorch#104501 is merged"; the torch in
this image ships that class (``torch.distributions.InverseGamma``, same ``(concentration, rate,
validate_args)`` signature), so the name is re-exported rather than restated. It is host-side only:
the engine has no inverse-gamma log-density (SURVEY.md §2 row 4 marks it out of scope - no
BASELINE.json configuration uses it), so a model that scores it raises ``NotImplementedError``.
"""
from torch.distributions import InverseGamma

__all__ = ["InverseGamma"]
