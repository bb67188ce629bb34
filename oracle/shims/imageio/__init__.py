"""Empty stand-in: robotarium_gym/utilities/misc.py:8 imports imageio for gif export only (oracle only)."""
