"""``python -m src.mnist --train/--sample`` — the reference's MNIST entry point, B200-native.

Thin alias of tinydiffusionmodels_b200.mnist so that code written against the reference
(``from src.mnist import SimpleUNet, q_sample, p_sample, sample, train``) keeps working.
"""
from tinydiffusionmodels_b200.mnist import (  # noqa: F401
    ResidualBlock, SimpleUNet, alphas, alphas_cumprod, betas, eval_mode, linear_beta_schedule, main, p_sample,
    q_sample, sample, sample_images, sample_loop, sqrt_alphas_cumprod,
    sqrt_one_minus_alphas_cumprod, timesteps, train,
)
from tinydiffusionmodels_b200.utils import (  # noqa: F401  (names the reference module also carries)
    get_samples_dir, get_vertex_checkpoint_path, load_checkpoint, save_checkpoint, save_samples,
)

if __name__ == "__main__":
    main()
