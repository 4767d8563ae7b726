#!/bin/bash
# Golden images from the reference's own Go binary, for tests/test_reference_go_output.py::test_go_goldens_when_present.
# Needs a Go toolchain and a checkout of igrega348/xray_projection_render (neither exists in the build image or on the GPU
# box, which is why tests/golden/go/ holds only the case list).  usage: tools/make_go_goldens.sh /path/to/xray_projection_render
# Every case of tests/golden/go/cases.json is rendered with the reference CLI (main.go:638-851) and its PNG frames are
# copied to tests/golden/go/<name>_NNN.png, untouched.
set -euo pipefail
ref=${1:?path to the reference checkout}
here=$(cd "$(dirname "$0")/.." && pwd)
out=$here/tests/golden/go
tmp=$(mktemp -d)
(cd "$ref" && go build -o "$tmp/xray_render" .)
python3 - "$out/cases.json" <<'PY' > "$tmp/cases.tsv"
import json, sys
for c in json.load(open(sys.argv[1]))["cases"]:
    print("\t".join(str(v) for v in (c["name"], c["input"], c["deformation"] or "-", c["resolution"], c["R"], c["fov"], c["ds"], c["integration"],
                                     c["flat_field"], c["density_multiplier"], ",".join(map(str, c["azimuthal"])), ",".join(map(str, c["polar"])))))
PY
while IFS=$'\t' read -r name input deform res R fov ds integ ff dm az polar; do
    args=(--input "examples/$input" --output_dir "$tmp/$name" --fname_pattern "${name}_%03d.png" --resolution "$res" --R "$R" --fov "$fov"
          --ds "$ds" --integration "$integ" --flat_field "$ff" --density_multiplier "$dm" --azimuthal_angles "$az" --polar_angles "$polar"
          --transforms_file "$tmp/$name/transforms.json" --text_progress)
    [ "$deform" != "-" ] && args+=(--deformation_file "examples/$deform")
    mkdir -p "$tmp/$name"
    (cd "$ref" && "$tmp/xray_render" "${args[@]}" > "$tmp/$name/log.txt" 2>&1) || { cat "$tmp/$name/log.txt"; exit 1; }
    cp "$tmp/$name"/${name}_*.png "$out/"
    echo "$name: $(ls "$tmp/$name"/${name}_*.png | wc -l) frame(s)"
done < "$tmp/cases.tsv"
echo "goldens in $out -- run: python -m pytest tests/test_reference_go_output.py -q"
