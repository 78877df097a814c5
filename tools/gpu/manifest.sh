#!/bin/bash
# Guards a queued gpurun call against a half-edited tree: `make` records a checksum of the
# sources + built libraries at a consistent point; `check` (first line of a GPU script)
# refuses to spend GPU minutes if the snapshot differs.
cd "$(dirname "$0")/../.."
sum() { find cartographer_b200 include oracle tests benchmarks bench.py __graft_entry__.py \
          -type f \( -name '*.py' -o -name '*.cu' -o -name '*.cuh' -o -name '*.h' -o -name '*.cc' \
          -o -name 'Makefile' -o -name 'libcsm_b200.so' -o -name 'liboracle.so' \) \
          -not -path '*/__pycache__/*' | sort | xargs sha1sum | sha1sum | cut -d' ' -f1; }
case "$1" in
  make) sum > .gpurun_manifest; cat .gpurun_manifest ;;
  check) [ "$(sum)" = "$(cat .gpurun_manifest 2>/dev/null)" ] || { echo "MANIFEST MISMATCH: snapshot taken mid-edit"; exit 9; } ;;
esac
