#!/bin/bash
# Where the two roles of k_fwd_stream spend their time (clock64 accumulators per phase, CTA 0):
# builds a traced copy of the library and runs one C2 forward.  tools only.
set -e
cd "$(dirname "$0")/.."
lib=$(bash tools/build_variant.sh trace "-DVEON_FWD_TRACE $EXTRA" | tail -1)
VEON_LIB=$lib python tools/fwd_trace.py "$@"
