"""TEST INFRASTRUCTURE -- CPU oracle of the reference path (see oracle/quatro_oracle.cpp header).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
from .oracle_lib import Oracle  # noqa: F401
