"""Empty stand-in: robotarium_gym/utilities/misc.py:10 imports tensorflow for eval logging only (oracle only)."""
