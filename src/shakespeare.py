"""``python -m src.shakespeare --sample/--guided_sample`` — the reference's text entry point,
B200-native (alias of tinydiffusionmodels_b200.shakespeare)."""
from tinydiffusionmodels_b200.shakespeare import (  # noqa: F401
    LearnedEmbedding, LearnedRounding, T, TinyTransformer, alphas, alphas_cumprod, betas, guided_generate,
    linear_beta_schedule, main, p_sample, q_sample, round_to_tokens, sample, sample_diffusion_embeddings,
    sqrt_alphas_cumprod, sqrt_one_minus_alphas_cumprod,
)

if __name__ == "__main__":
    main()
