#!/bin/sh
# TEST / BENCHMARK INFRASTRUCTURE.  Stages the reference's own hot-path modules (unmodified, byte for byte) from
# the read-only checkout into oracle/_ref/, which is git-ignored (never part of this repository's history) but
# travels to the GPU box with the built artefacts, so that `bench.py --impl reference` can time the reference's
# own CPU path (RawData + ChunkRawRecords) on the GPU box's host cores.  oracle/ref_loader.py executes the files
# from there exactly as it executes them from /root/reference here.  Nothing under wfsim_b200/ reads oracle/_ref.
set -e
SRC="${WFSIM_REFERENCE_ROOT:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref"
[ -f "$SRC/wfsim/core/pulse.py" ] || { echo "make_ref: no reference checkout at $SRC (nothing staged)"; exit 0; }
rm -rf "$DST"
mkdir -p "$DST/wfsim/core" "$DST/files"
for f in units.py utils.py load_resource.py strax_interface.py; do cp "$SRC/wfsim/$f" "$DST/wfsim/$f"; done
for f in pulse.py s1.py s2.py afterpulse.py rawdata.py; do cp "$SRC/wfsim/core/$f" "$DST/wfsim/core/$f"; done
cp "$SRC/files/XENONnT_spe_distributions_single_channel.csv" "$DST/files/"
(cd "$SRC" && sha256sum wfsim/units.py wfsim/utils.py wfsim/load_resource.py wfsim/strax_interface.py \
    wfsim/core/pulse.py wfsim/core/s1.py wfsim/core/s2.py wfsim/core/afterpulse.py wfsim/core/rawdata.py) > "$DST/SHA256SUMS"
echo "make_ref: staged $(ls "$DST/wfsim" "$DST/wfsim/core" | grep -c '\.py$') modules into $DST"
