#!/bin/bash
# one-minute sanity of a rebuilt library: the one-to-many tests, smoke()
timeout 80 python -m pytest tests/test_gpu_join.py -m gpu -q -x --timeout 60 -k "one_to_many or duplicates or heavy or golden or hash_build_dwarf or groupby" 2>&1 | tail -3 | cut -c1-300
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
