"""pylbm_b200: B200-native lattice Boltzmann time-step engine behind the pylbm API."""
