#!/bin/bash
# Pins the oracle (and through it the CUDA path) to the reference itself. Needs a Rust toolchain, which the build image
# of this repository does not have: run it wherever `cargo` exists.
#
#   tools/pin_against_dryv.sh /path/to/dryv        # a checkout of Stuff7/dryv
#
# For every tests/golden/pin/*.mp4 it runs `dryv <file>` (src/main.rs: Video::open decodes the first sample and writes its
# frame to ./temp/yuv_frame, src/video/decoder.rs:141-143) and compares the SHA-256 of that file with the digest the
# oracle predicts (tests/golden/pin/SHA256SUMS, written by tools/make_pin_kit.py). The files are ISO BMFF with an "isom"
# major brand (src/video/decoder.rs:45-54), ftyp / moov / mdat at the top level (atom/root.rs:17-52), one video track
# (hdlr "vide") with tkhd, mdia{mdhd, hdlr, minf{vmhd, dinf, stbl{stsd{avc1{avcC}}, stts, stss, stsc, stsz, stco}}}: the boxes
# Video::open walks (src/video/mod.rs:44-104) and SampleIter needs (sample/mod.rs:74-110).
set -u
DRYV=${1:?usage: tools/pin_against_dryv.sh /path/to/dryv}
KIT="$(cd "$(dirname "$0")/.." && pwd)/tests/golden/pin"
cd "$DRYV" || exit 2
cargo build --release || exit 2
mkdir -p temp/slice    # the decoder dumps every slice to temp/slice/<n> and expects the directory (decoder.rs:127-139)
fail=0
while read -r want name _; do
  case "$want" in \#*|"") continue ;; esac
  rm -f temp/yuv_frame
  ./target/release/dryv "$KIT/$name" > /dev/null 2>&1
  if [ ! -f temp/yuv_frame ]; then echo "FAIL $name: dryv wrote no temp/yuv_frame"; fail=1; continue; fi
  got=$(sha256sum temp/yuv_frame | cut -d' ' -f1)
  if [ "$got" = "$want" ]; then echo "ok   $name"; else echo "FAIL $name: dryv $got, oracle $want"; fail=1; fi
done < "$KIT/SHA256SUMS"
exit $fail
