"""dsopp_b200 -- B200-native photometric bundle-adjustment hot path for DSOPP."""
