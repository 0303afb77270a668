"""mininf_b200: B200-native ELBO engine behind the mininf API."""
