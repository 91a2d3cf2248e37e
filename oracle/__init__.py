"""CPU oracle for the HGI hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (rustyhgi_b200/) never does.

`oracle.c` wraps oracle/libhgi_oracle.so (hgi_oracle.c, the literal C restatement of the
reference).  `oracle.pyref` is an independent numpy restatement used to cross-check it.
"""
from . import c, pyref  # noqa: F401
