#!/bin/bash
# build the library in-tree and fail loudly (used before every gpurun call)
cd "$(dirname "$0")/.." && python -c "
import sys; sys.path.insert(0,'.')
import __graft_entry__ as g; g.build()" > /tmp/build.log 2>&1 || { tail -30 /tmp/build.log; echo BUILD FAILED; exit 1; }
tail -2 /tmp/build.log
