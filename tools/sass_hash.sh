#!/bin/bash
# md5 of the SASS of the benched kernel instantiation (k_push3<strict FP, no stats, fast div/sqrt>) in the in-tree library:
# `tools/sass_hash.sh` prints it, `tools/sass_hash.sh record` writes profiles/k_push3_sass.md5 (tests/test_abi.py compares).
cd "$(dirname "$0")/.."
H=$(cuobjdump -sass -fun '_ZN4cpic7k_push3ILb0ELb0ELb1EEEvNS_9Push3ArgsEf' cabanapic_b200/libcabanapic_b200.so 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's#/\* 0x[0-9a-f]* \*/##' | md5sum | cut -d' ' -f1)
if [ "$1" = record ]; then echo "$H" > profiles/k_push3_sass.md5; fi
echo "$H"
