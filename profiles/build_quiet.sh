#!/bin/bash
# build the library, printing only the first lines of each diagnostic (macro expansions make nvcc's messages huge)
cd "$(dirname "$0")/.." && python __graft_entry__.py --force 2>&1 | grep -E "error|warning|built|Error" | cut -c1-400 | head -${1:-40}
