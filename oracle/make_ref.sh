#!/bin/bash
# Recipe: stage the UNMODIFIED reference's own Python implementation of the hot path (and the loss /
# optimiser code its train step runs) under oracle/_ref/, from the sources where they lie in the reference
# checkout.  oracle/_ref/ is git-ignored (never part of this repo's history) but NOT gpurun-ignored, so it
# travels to the GPU box and `bench.py --impl reference[-gpu]` can time the real reference there.
# Nothing under oracle/_ref is imported by the product (hs-pose_b200/); see oracle/ref_loader.py.
set -e
REF=${HSPOSE_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref"
[ -d "$REF/network/fs_net_repo" ] || { echo "reference tree not found at $REF — skipping"; exit 0; }
rm -rf "$DST"
mkdir -p "$DST"
for d in network losses config engine; do
  mkdir -p "$DST/$d"
done
cp -r "$REF/network/fs_net_repo" "$DST/network/"
cp "$REF/network/HSPose.py" "$DST/network/"
mkdir -p "$DST/network/point_sample" && cp "$REF/network/point_sample/"*.py "$DST/network/point_sample/" 2>/dev/null || true
cp "$REF/losses/"*.py "$DST/losses/"
cp "$REF/config/"*.py "$DST/config/"
cp "$REF/engine/organize_loss.py" "$DST/engine/"
mkdir -p "$DST/tools/torch_utils/solver" "$DST/datasets" "$DST/tools/lynne_lib"
for f in rot_utils plane_utils geom_utils training_utils solver_utils logger; do cp "$REF/tools/$f.py" "$DST/tools/"; done
cp "$REF/tools/torch_utils/solver/"*.py "$DST/tools/torch_utils/solver/"
[ -f "$REF/tools/torch_utils/__init__.py" ] && cp "$REF/tools/torch_utils/__init__.py" "$DST/tools/torch_utils/" || true
cp "$REF/datasets/data_augmentation.py" "$DST/datasets/"
cp "$REF/tools/lynne_lib/"*.py "$DST/tools/lynne_lib/" 2>/dev/null || true
for d in "$DST" "$DST/network" "$DST/losses" "$DST/config" "$DST/engine" "$DST/tools" "$DST/datasets" \
         "$DST/tools/torch_utils" "$DST/tools/torch_utils/solver" "$DST/tools/lynne_lib" "$DST/network/point_sample"; do
  [ -f "$d/__init__.py" ] || : > "$d/__init__.py"
done
rm -f "$DST/__init__.py"
( cd "$REF" && git rev-parse HEAD 2>/dev/null || echo unknown ) > "$DST/REFERENCE_COMMIT"
echo "staged reference hot path under $DST ($(find "$DST" -name '*.py' | wc -l) files)"
