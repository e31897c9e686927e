"""B200-native differentiable MLS-MPM engine behind DiffSkill's plb.engine API."""
__version__ = "0.1.0"
