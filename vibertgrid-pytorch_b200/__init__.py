"""vibertgrid_pytorch_b200 -- B200-native ViBERTgrid joint forward."""
__version__ = "0.1.0"
