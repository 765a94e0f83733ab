#!/usr/bin/env bash
# oracle/build_ref.sh — build the REAL reference (kcftools, Java 17 + Maven) into oracle/_ref/ when a JDK is present, and
# emit reference-authored fixtures for the getVariations path.  The reference is Java with un-vendored Maven dependencies
# (picocli 4.7.6, jgrapht-core 1.5.2, commons-lang3, jetbrains annotations: pom.xml:20-55); this image has no JDK, no
# Maven and no network, so on this box the script records that and exits 0 — the recipe exists so that the day a JDK and
# the dependency jars are reachable, parity can be pinned against the reference itself.
#
#   oracle/_ref/kcftools.jar                 the reference, compiled from /root/reference/src/main/java as it lies there
#   oracle/_ref/STATUS                       what happened (always written)
#   tests/fixtures_reference/*.kcf           getVariations output of the reference on tests/fixtures_oracle/small.npz's inputs
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${KCF_REFERENCE_DIR:-/root/reference}"
OUT="$ROOT/oracle/_ref"
mkdir -p "$OUT"
status() { echo "$(date -u +%FT%TZ) $*" | tee "$OUT/STATUS"; }

JAVAC="$(command -v javac || true)"
JAVA="$(command -v java || true)"
if [ -z "$JAVAC" ] || [ -z "$JAVA" ]; then
    status "no JDK (javac/java not on PATH): reference not built; bench.py --impl reference times the C port (oracle/kcf_oracle.c)"
    exit 0
fi
if [ ! -d "$REF/src/main/java" ]; then
    status "reference sources not found under $REF: nothing to build"
    exit 0
fi
# dependency jars: a local Maven repository or a directory of jars given by KCF_REF_JARS (no network here)
CP=""
for d in "${KCF_REF_JARS:-}" "$HOME/.m2/repository"; do
    [ -n "$d" ] && [ -d "$d" ] && CP="$CP:$(find "$d" -name '*.jar' \( -name 'picocli-*' -o -name 'jgrapht-core-*' -o -name 'commons-lang3-*' -o -name 'annotations-*' -o -name 'jheaps-*' \) | paste -sd: -)"
done
CLS="$OUT/classes"
rm -rf "$CLS" && mkdir -p "$CLS"
if ! "$JAVAC" -nowarn -d "$CLS" -cp ".$CP" $(find "$REF/src/main/java" -name '*.java') 2> "$OUT/javac.log"; then
    status "javac failed (dependency jars missing? see oracle/_ref/javac.log; set KCF_REF_JARS): reference not built"
    exit 0
fi
# version.properties is filled by Maven resource filtering (pom.xml:58-63): restate it
mkdir -p "$CLS" && printf 'version=0.4.0\n' > "$CLS/version.properties"
( cd "$CLS" && jar cfe "$OUT/kcftools.jar" nl.wur.bis.kcftools.Main.KCFTOOLS . ) || { status "jar failed"; exit 0; }
# reference-authored fixtures: run getVariations on the small case's files
FX="$ROOT/tests/fixtures_reference"
mkdir -p "$FX"
python3 "$ROOT/tools/make_fixture.py" --write-files "$FX/small" > /dev/null 2>&1 || true
if [ -f "$FX/small/ref.fa" ]; then
    for mode in window gene transcript; do
        args="-f $mode"; [ "$mode" = window ] && args="$args -w 5000" || args="$args -g $FX/small/ann.gtf"
        "$JAVA" -cp "$OUT/kcftools.jar$CP" nl.wur.bis.kcftools.Main.KCFTOOLS getVariations -r "$FX/small/ref.fa" -k "$FX/small/sample" \
            -o "$FX/small_$mode.kcf" -s small $args > "$FX/small_$mode.log" 2>&1 || true
    done
    status "reference built: $OUT/kcftools.jar; fixtures under tests/fixtures_reference/ (compare with tests/test_fixture_small.py)"
else
    status "reference built: $OUT/kcftools.jar; fixture inputs not written (tools/make_fixture.py --write-files)"
fi
exit 0
