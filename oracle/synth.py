"""Seeded synthetic inputs / weights (clips, priors, state dicts, metric pairs, evaluation trees).  The generators live in the
package (``iip_uavsal_saliency_b200.synth``) because ``bench.py``'s product arm needs inputs without touching ``oracle/``;
this module re-exports them for the tests, the golden-vector scripts and the tools, which import them from here."""
from iip_uavsal_saliency_b200.synth import *          # noqa: F401,F403
from iip_uavsal_saliency_b200.synth import _init_rule, _upsample_linear, load_key_table          # noqa: F401
