"""``python -m src.shakespeare --train/--sample/--guided_sample`` — the reference's text entry point,
B200-native (alias of tinydiffusionmodels_b200.shakespeare)."""
from tinydiffusionmodels_b200.shakespeare import (  # noqa: F401
    HF_TOKEN, LearnedEmbedding, LearnedRounding, T, TinyTransformer, alphas, alphas_cumprod, betas,
    cosine_warmup_factor, dynamic_rounding_weight_schedule, get_cosine_schedule_with_warmup, guided_generate,
    linear_beta_schedule, load_text_dataset, main, p_sample, q_sample, round_to_tokens, sample,
    sample_diffusion_embeddings, sqrt_alphas_cumprod, sqrt_one_minus_alphas_cumprod, tokenize_corpus, train,
)
from tinydiffusionmodels_b200.utils import (  # noqa: F401  (names the reference module also carries)
    get_samples_dir, get_vertex_checkpoint_path, load_checkpoint, save_checkpoint, save_samples,
)

if __name__ == "__main__":
    main()
