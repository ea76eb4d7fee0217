timeout 600 python scratch/e2e_steps.py c2 2>&1 | tail -6
