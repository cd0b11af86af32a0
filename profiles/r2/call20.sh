python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "known_answer or residency" 2>&1 | tail -2
