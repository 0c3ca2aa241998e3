#!/bin/bash
# Fingerprint of the DEVICE code of every object of the product library: md5 of the SASS instruction stream with addresses, encodings
# and (path-dependent) anonymous-namespace symbol names stripped.  Host-only edits must leave it unchanged; the fingerprint of the
# build that last passed `pytest -m gpu` on a B200 is kept in profiles/r1_sass_fingerprint.txt.
#   tools/sass_fingerprint.sh [build-dir]
B=${1:-$(dirname $0)/../kintinuous_b200/csrc/build}
for o in $B/*.o; do
    h=$(cuobjdump -sass $o | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/\/\*.*//' | md5sum | cut -c1-16)
    echo "$(basename $o) $h"
done
