"""Development timing: one Nk=1000 solve (eta from argv, default 0.51), twice."""
import os, sys
sys.path.insert(0, os.path.dirname(__file__))
from quick_bench import run
eta = float(sys.argv[1]) if len(sys.argv) > 1 else 0.51
run(1000, eta); run(1000, eta)
