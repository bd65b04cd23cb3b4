"""Drop-in model classes (same names, constructor and forward signatures, attributes and state_dict keys as the
reference's ``models_fer_vit`` package); forward/backward run as native sm_100a plans."""
from .latent_vit import LatentViT
from .latent_vit_v2 import LatentViTv2
from .hybrid_latent_vit import HybridLatentViT, AdapterModule, create_hybrid_latent_vit, RECOMMENDED_STRATEGIES
from .image_vit import ImageViT, PatchEmbedding, create_vit_tiny, create_vit_small, create_vit_base
from .latent_decomposer import LatentDecomposer
from .expression_aware_vit import ExpressionAwareViT

__all__ = ["LatentViT", "LatentViTv2", "HybridLatentViT", "AdapterModule", "create_hybrid_latent_vit",
           "RECOMMENDED_STRATEGIES", "ImageViT", "PatchEmbedding", "create_vit_tiny", "create_vit_small",
           "create_vit_base", "LatentDecomposer", "ExpressionAwareViT"]
