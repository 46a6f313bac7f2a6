python tests/wgrad_prof.py 2 128 16 2>&1 | tail -10
python tests/wgrad_prof.py 2 64 32 2>&1 | tail -10
python tests/wgrad_prof.py 2 32 64 2>&1 | tail -10
python tests/wgrad_prof.py 2 16 128 2>&1 | tail -10
