#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -n 4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
