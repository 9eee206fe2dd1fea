"""Stand-in for `simplejson` (TEST INFRASTRUCTURE ONLY): the stdlib json has the same API
for what dispersion.pyx:483-549 uses."""
from json import *  # noqa
