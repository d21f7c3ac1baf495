#!/bin/bash
# one-minute sanity of a rebuilt library: the one-to-many tests, smoke()
timeout 50 python -m pytest tests/test_gpu_join.py -m gpu -q -x --timeout 40 -k "${1:-one_to_many or duplicates or heavy or golden or hash_build_dwarf or groupby}" 2>&1 | tail -8 | cut -c1-400
[ -n "$1" ] || timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
