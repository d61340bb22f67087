#!/bin/bash
# Build the reference's post-processing tool (oracle/_ref/post_<cfg>) for the golden configurations that have walls.
# usage: bash oracle/refbuild/build_post_all.sh
HERE="$(cd "$(dirname "$0")" && pwd)"
source "$HERE/configs.sh"
for name in chan_s3v2 chan_s2v2; do
  [ -x "$HERE/../_ref/post_$name" ] || bash "$HERE/build_ref_post.sh" "$name" ${CFG[$name]} || exit 1
done
