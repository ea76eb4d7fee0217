"""vivit_b200 -- B200-native low-rank GGN hot path with ViViT's BackPACK-facing API."""

__version__ = "0.1.0"
