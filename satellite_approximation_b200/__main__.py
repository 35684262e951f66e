"""`python -m satellite_approximation_b200 laplace_main|poisson_main|fill_folder ...` = drivers.main (the reference's
executables restated on the GPU path; see drivers.py)."""
import sys

from .drivers import main

sys.exit(main())
