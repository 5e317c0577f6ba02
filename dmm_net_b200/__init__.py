"""dmm_net_b200 -- B200-native (sm_100a) implementation of DMM-Net's differentiable mask-matching layer."""
__version__ = "0.2.0"
