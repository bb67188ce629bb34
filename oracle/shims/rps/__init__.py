"""Stand-in for robotarium_python_simulator (`rps`), pinned by the reference README to commit
6bb184e, ORACLE ONLY (test infrastructure).

rps is NOT in /root/reference and cannot be installed here.  The modules below restate the slice
of rps that MARBLER's hot path calls: the unicycle
simulator step/validate, the SI position controller, the SI<->unicycle maps, the single-integrator
barrier certificates and the grid spawn sampler.  PARITY UNPINNED at this boundary: every constant
that might differ at 6bb184e is a named module-level constant.
"""
