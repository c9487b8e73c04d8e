"""vyomai_b200 — B200-native (sm_100a) transformer-block hot path behind VyomAI's module API."""
__version__ = "0.1.0"
