"""Development timing: Nk=1000 (eta=0.51, 16 137 evaluations), Nk=500 and Nk=100 single solves; prints cycles per evaluation."""
import os, sys
sys.path.insert(0, os.path.dirname(__file__))
from quick_bench import run
for Nk, eta in ((1000, 0.51), (1000, 0.51), (500, 0.51), (100, 0.51)):
    run(Nk, eta)
