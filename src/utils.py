"""The reference's ``src/utils.py`` surface (checkpoint / sample I/O), plus — as BASELINE.json's
north_star asks — the noise-schedule and q_sample helpers re-exported from one place."""
from tinydiffusionmodels_b200.utils import (  # noqa: F401
    download_from_gcs, get_samples_dir, get_vertex_checkpoint_path, is_gcs_path, load_checkpoint,
    parse_gcs_path, save_checkpoint, save_samples, storage, upload_to_gcs,
)
from tinydiffusionmodels_b200.schedule import linear_beta_schedule, make_schedule, schedule_on  # noqa: F401
from tinydiffusionmodels_b200.mnist import q_sample  # noqa: F401
