#!/bin/bash
# usage: tools/sass_of.sh <lib.so> [mangled kernel name]  ->  SASS of one kernel (default: the stereo bit-exact float granule kernel), one instruction per line
k=${2:-_ZN3l3b17l3_granule_kernelILi2ELi4ELb0ELb0ELb0ELb0EEEvNS_11BatchParamsEPKNS_4TileEj}
cuobjdump -sass -fun "$k" "$1" | grep -E '^\s+/\*[0-9a-f]{4,6}\*/' | sed -E 's/^\s+\/\*([0-9a-f]+)\*\/\s+/\1 /; s/\s*\/\*.*$//'
