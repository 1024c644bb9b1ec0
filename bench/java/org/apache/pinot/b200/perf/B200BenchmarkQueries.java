/**
 * JVM timing harness for the reference's OWN Java path on bench.py's workload -- the "reference JVM path" comparator
 * BASELINE.md section 2 asks for.  SOURCE ONLY in this repository: the build image has no JDK / Maven / jars, so it has
 * never been compiled here; bench/java/run_jvm_baseline.sh compiles and runs it when a JDK and the reference's
 * pinot-core + pinot-perf test classpath are present, and says "no JDK" otherwise.
 *
 * What it times (same table and SQL as bench.py):
 *   table     SEGMENTS x ROWS rows, 8 dict-encoded INT columns c0..c7, cardinalities {10,100,1000,10000,65536,100000,1e6,1e6},
 *             dictId(doc) = mix64(seed + doc * 0x9E3779B97F4A7C15) % cardinality, value = base + step * dictId
 *             (the generator of pinot_b200/csrc/pb200_synth.cu and oracle/pinot_oracle.cpp po_synth_fwd)
 *   headline  SELECT SUM(c5), COUNT(*) FROM benchTable WHERE c6 > K GROUP BY c3 LIMIT 100000
 *   c2        SELECT SUM(c5), COUNT(*) FROM benchTable WHERE c3 BETWEEN lo AND hi AND c6 > K
 *   (a) operatorOnly: getOperator(sql).nextBlock() per segment on ONE thread each (the per-segment seam the device
 *       library replaces: BaseQueriesTest.getOperator, pinot-core/src/test/.../queries/BaseQueriesTest.java:97-102)
 *   (b) instancePlan: PLAN_MAKER.makeInstancePlan(segments, queryContext, executor, null).execute() with
 *       maxExecutionThreads = #segments (BaseQueriesTest.java:218-231) -- the server-level block bench.py's step delivers
 * JMH AverageTime like pinot-perf's BenchmarkQueries (pinot-perf/src/main/java/org/apache/pinot/perf/BenchmarkQueries.java:68-74).
 * Prints Runtime.availableProcessors() so the number can be quoted with its core count.
 */
package org.apache.pinot.b200.perf;

import java.io.File;
import java.util.ArrayList;
import java.util.List;
import java.util.concurrent.ExecutorService;
import java.util.concurrent.Executors;
import java.util.concurrent.TimeUnit;
import org.apache.commons.io.FileUtils;
import org.apache.pinot.common.request.PinotQuery;
import org.apache.pinot.core.common.Operator;
import org.apache.pinot.core.operator.blocks.results.BaseResultsBlock;
import org.apache.pinot.core.plan.Plan;
import org.apache.pinot.core.plan.maker.InstancePlanMakerImplV2;
import org.apache.pinot.core.plan.maker.PlanMaker;
import org.apache.pinot.core.query.request.context.QueryContext;
import org.apache.pinot.core.query.request.context.utils.QueryContextConverterUtils;
import org.apache.pinot.segment.local.indexsegment.immutable.ImmutableSegmentLoader;
import org.apache.pinot.segment.local.segment.creator.impl.SegmentIndexCreationDriverImpl;
import org.apache.pinot.segment.local.segment.readers.GenericRowRecordReader;
import org.apache.pinot.segment.spi.IndexSegment;
import org.apache.pinot.segment.spi.SegmentContext;
import org.apache.pinot.segment.spi.creator.SegmentGeneratorConfig;
import org.apache.pinot.spi.config.table.TableConfig;
import org.apache.pinot.spi.config.table.TableType;
import org.apache.pinot.spi.data.FieldSpec;
import org.apache.pinot.spi.data.Schema;
import org.apache.pinot.spi.data.readers.GenericRow;
import org.apache.pinot.spi.utils.ReadMode;
import org.apache.pinot.spi.utils.builder.TableConfigBuilder;
import org.apache.pinot.sql.parsers.CalciteSqlParser;
import org.openjdk.jmh.annotations.Benchmark;
import org.openjdk.jmh.annotations.BenchmarkMode;
import org.openjdk.jmh.annotations.Fork;
import org.openjdk.jmh.annotations.Measurement;
import org.openjdk.jmh.annotations.Mode;
import org.openjdk.jmh.annotations.OutputTimeUnit;
import org.openjdk.jmh.annotations.Param;
import org.openjdk.jmh.annotations.Scope;
import org.openjdk.jmh.annotations.Setup;
import org.openjdk.jmh.annotations.State;
import org.openjdk.jmh.annotations.TearDown;
import org.openjdk.jmh.annotations.Warmup;
import org.openjdk.jmh.infra.Blackhole;
import org.openjdk.jmh.runner.Runner;
import org.openjdk.jmh.runner.options.OptionsBuilder;


@BenchmarkMode(Mode.AverageTime)
@OutputTimeUnit(TimeUnit.MILLISECONDS)
@Fork(1)
@Warmup(iterations = 3, time = 5)
@Measurement(iterations = 5, time = 5)
@State(Scope.Benchmark)
public class B200BenchmarkQueries {
  private static final String TABLE = "benchTable";
  private static final int[] CARDS = {10, 100, 1_000, 10_000, 65_536, 100_000, 1_000_000, 1_000_000};
  private static final int[] VALUE_STEP = {1, 1, 1, 3, 1, 7, 2, 2};
  private static final int[] VALUE_BASE = {0, 0, 0, 5, 0, 11, 1, 1};
  private static final PlanMaker PLAN_MAKER = new InstancePlanMakerImplV2();

  @Param({"8"})
  public int _segments;
  @Param({"100000000"})
  public int _rows;
  /** group-by (bench.py headline, 10 % of the rows) or c2 (2-predicate range filter, 25 %) */
  @Param({"groupby", "c2"})
  public String _query;

  private final List<IndexSegment> _indexSegments = new ArrayList<>();
  private File _indexDir;
  private ExecutorService _executor;
  private QueryContext _queryContext;

  private static long mix64(long z) {
    z = (z ^ (z >>> 30)) * 0xBF58476D1CE4E5B9L;
    z = (z ^ (z >>> 27)) * 0x94D049BB133111EBL;
    return z ^ (z >>> 31);
  }

  static String sql(String which) {
    if (which.equals("groupby")) {
      int kId = (int) Math.round(CARDS[6] * 0.9) - 1;
      return "SELECT SUM(c5), COUNT(*) FROM " + TABLE + " WHERE c6 > " + (VALUE_BASE[6] + VALUE_STEP[6] * kId)
          + " GROUP BY c3 LIMIT 100000";
    }
    double f = Math.sqrt(0.25);
    int loId = (int) (CARDS[3] * (1 - f) / 2);
    int hiId = loId + (int) Math.round(CARDS[3] * f) - 1;
    int kId = (int) Math.round(CARDS[6] * (1 - f)) - 1;
    return "SELECT SUM(c5), COUNT(*) FROM " + TABLE + " WHERE c3 BETWEEN " + (VALUE_BASE[3] + VALUE_STEP[3] * loId) + " AND "
        + (VALUE_BASE[3] + VALUE_STEP[3] * hiId) + " AND c6 > " + (VALUE_BASE[6] + VALUE_STEP[6] * kId);
  }

  @Setup
  public void setUp()
      throws Exception {
    System.out.println("availableProcessors = " + Runtime.getRuntime().availableProcessors());
    _indexDir = new File(FileUtils.getTempDirectory(), "B200BenchmarkQueries");
    FileUtils.deleteQuietly(_indexDir);
    Schema.SchemaBuilder sb = new Schema.SchemaBuilder().setSchemaName(TABLE);
    for (int c = 0; c < 8; c++) {
      sb.addSingleValueDimension("c" + c, FieldSpec.DataType.INT);
    }
    Schema schema = sb.build();
    TableConfig tableConfig = new TableConfigBuilder(TableType.OFFLINE).setTableName(TABLE).build();
    for (int s = 0; s < _segments; s++) {
      final int seg = s;
      // rows are produced lazily: a 100 M-row segment never sits in a List<GenericRow>
      List<GenericRow> rows = new java.util.AbstractList<GenericRow>() {
        @Override
        public GenericRow get(int doc) {
          GenericRow row = new GenericRow();
          for (int c = 0; c < 8; c++) {
            long seed = 1000L + 131L * seg + c;   // rank 0 of bench.py's column_specs
            int dictId = (int) Long.remainderUnsigned(mix64(seed + doc * 0x9E3779B97F4A7C15L), CARDS[c]);
            row.putValue("c" + c, VALUE_BASE[c] + VALUE_STEP[c] * dictId);
          }
          return row;
        }

        @Override
        public int size() {
          return _rows;
        }
      };
      SegmentGeneratorConfig config = new SegmentGeneratorConfig(tableConfig, schema);
      config.setOutDir(_indexDir.getPath());
      config.setTableName(TABLE);
      config.setSegmentName("r0s" + s);
      SegmentIndexCreationDriverImpl driver = new SegmentIndexCreationDriverImpl();
      driver.init(config, new GenericRowRecordReader(rows));
      driver.build();
      _indexSegments.add(ImmutableSegmentLoader.load(new File(_indexDir, "r0s" + s), ReadMode.mmap));
    }
    _executor = Executors.newFixedThreadPool(Math.max(2, _segments));
    PinotQuery pinotQuery = CalciteSqlParser.compileToPinotQuery(sql(_query));
    _queryContext = QueryContextConverterUtils.getQueryContext(pinotQuery);
    _queryContext.setEndTimeMs(Long.MAX_VALUE);
    _queryContext.setMaxExecutionThreads(_segments);
  }

  @TearDown
  public void tearDown() {
    for (IndexSegment s : _indexSegments) {
      s.destroy();
    }
    _executor.shutdownNow();
    FileUtils.deleteQuietly(_indexDir);
  }

  /** (a) the per-segment operator chain, one segment after the other on the calling thread */
  @Benchmark
  public void operatorOnly(Blackhole bh) {
    for (IndexSegment segment : _indexSegments) {
      Operator<?> op = PLAN_MAKER.makeSegmentPlanNode(new SegmentContext(segment), _queryContext).run();
      bh.consume(op.nextBlock());
    }
  }

  /** (b) the server-level plan: all segments in parallel + combine, what bench.py's step returns as ONE block */
  @Benchmark
  public void instancePlan(Blackhole bh) {
    List<SegmentContext> contexts = new ArrayList<>();
    for (IndexSegment segment : _indexSegments) {
      contexts.add(new SegmentContext(segment));
    }
    Plan plan = PLAN_MAKER.makeInstancePlan(contexts, _queryContext, _executor, null);
    BaseResultsBlock block = plan.execute().getResultsBlock();
    bh.consume(block);
  }

  public static void main(String[] args)
      throws Exception {
    new Runner(new OptionsBuilder().include(B200BenchmarkQueries.class.getSimpleName()).build()).run();
  }
}
