"""The reference's ``src/utils.py`` surface (checkpoint / sample I/O), plus — as BASELINE.json's
north_star asks — the noise-schedule and q_sample helpers re-exported from one place.

``src.utils`` IS ``tinydiffusionmodels_b200.utils`` (one module object under two names), so code and tests written
against the reference that reach into the module - ``patch('src.utils.download_from_gcs')``,
``src.utils.storage.Client`` - act on the implementation itself."""
import sys

import tinydiffusionmodels_b200.utils as _impl
from tinydiffusionmodels_b200.mnist import q_sample
from tinydiffusionmodels_b200.schedule import linear_beta_schedule, make_schedule, schedule_on

for _name, _obj in (("linear_beta_schedule", linear_beta_schedule), ("make_schedule", make_schedule),
                    ("schedule_on", schedule_on), ("q_sample", q_sample)):
    setattr(_impl, _name, _obj)
sys.modules[__name__] = _impl
