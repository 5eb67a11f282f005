"""TEST INFRASTRUCTURE ONLY -- a stand-in for the four `pyro-ppl` names the reference imports.

The reference pins pyro-ppl==1.6.0 (EPro-PnP-Det/requirements.txt:6; 1.4.0 for the 6DoF tree,
EPro-PnP-6DoF/README.md:35).  pyro is not installed here and cannot be (no network), so
`oracle/make_golden.py` puts this directory on sys.path *only while it imports the unmodified
reference from /root/reference* to generate golden vectors.  Nothing in the product imports it.

PARITY UNPINNED for the Student-t arithmetic: the reference holds no test or fixture for it and
pyro's own source cannot be diffed offline.  `distributions/__init__.py` restates the published
multivariate-t density/sampler and tests/test_oracle_cpu.py cross-checks the density against
scipy.stats.multivariate_t.
"""
__version__ = "0.0-shim"
