#!/bin/bash
# Runs the reference's OWN Java path on bench.py's workload (bench/java/.../B200BenchmarkQueries.java) -- when it can.
# Needs: a JDK 11+ (javac, java) and PINOT_CLASSPATH = the reference's pinot-core test classpath + pinot-perf (JMH), e.g.
#   (cd $REFERENCE && mvn -q -pl pinot-perf -am -DskipTests package dependency:build-classpath -Dmdep.outputFile=cp.txt)
#   export PINOT_CLASSPATH="$(cat $REFERENCE/pinot-perf/cp.txt):$REFERENCE/pinot-perf/target/classes"
# The build image of this repository has neither: the script then says so and exits 0 -- it never
# prints a fabricated number.
set -u
here="$(cd "$(dirname "$0")" && pwd)"
if ! command -v javac >/dev/null 2>&1 || ! command -v java >/dev/null 2>&1; then
  echo '{"impl": "reference-jvm", "unavailable": "no JDK (javac / java) on this machine: the Java reference path cannot be timed here"}'
  exit 0
fi
if [ -z "${PINOT_CLASSPATH:-}" ]; then
  echo '{"impl": "reference-jvm", "unavailable": "PINOT_CLASSPATH is not set (pinot-core + pinot-perf + JMH jars of the reference build)"}'
  exit 0
fi
out="${TMPDIR:-/tmp}/b200_jvm_baseline_classes"
mkdir -p "$out"
javac -cp "$PINOT_CLASSPATH" -d "$out" "$here/org/apache/pinot/b200/perf/B200BenchmarkQueries.java" || { echo '{"impl": "reference-jvm", "unavailable": "javac failed (see stderr)"}'; exit 0; }
echo "nproc=$(nproc) cpu=$(grep -m1 'model name' /proc/cpuinfo | cut -d: -f2-)"
exec java -Xms64g -Xmx64g -XX:MaxDirectMemorySize=64g -cp "$out:$PINOT_CLASSPATH" org.apache.pinot.b200.perf.B200BenchmarkQueries "$@"
