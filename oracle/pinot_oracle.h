/*
 * pinot_oracle.h -- C API of the CPU parity oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is a CPU restatement of the reference's (y-scope/pinot 1.3.0-SNAPSHOT) Java
 * algorithm for ONE path: per-segment scan -> filter -> project -> (group-by) aggregate.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.  The product (pinot_b200/)
 * never links, imports or calls anything in oracle/.
 *
 * Parity pinning: the oracle is checked (tests/test_oracle_golden.py) against
 *   - the reference's own known-answer assertions on test_data-sv.avro
 *     (pinot-core/src/test/java/org/apache/pinot/queries/InnerSegmentAggregationSingleValueQueriesTest.java:43-175,
 *      InterSegmentAggregationSingleValueQueriesTest.java:47-259), results AND ExecutionStatistics,
 *   - golden BYTES of reference-built index files (paddingOld.tar.gz forward index + dictionaries,
 *     data/startree/segment/star_tree_index).
 * RoaringBitmap (org.roaringbitmap:RoaringBitmap:1.3.0, pom.xml:799-801) is a third-party dependency absent from the
 * reference tree: its portable serialization is restated from the public RoaringFormatSpec; the reference holds no
 * golden bitmap bytes, so BYTE-level parity of Roaring serialization is unpinned (set-equality only).
 */
#ifndef PINOT_ORACLE_H_
#define PINOT_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- data types (stored types of FieldSpec.DataType) ---- */
enum { PO_INT = 0, PO_LONG = 1, PO_FLOAT = 2, PO_DOUBLE = 3, PO_STRING = 4 };

/* ---- column / segment description: pointers into Pinot-format index bytes ---- */
typedef struct {
  int32_t data_type;        /* PO_* */
  int32_t has_dictionary;   /* 1: fwd is a fixed-bit dictId stream (or sorted index); 0: raw fixed-byte chunk file */
  int32_t bits_per_value;   /* column.<c>.bitsPerElement */
  int32_t cardinality;      /* dictionary length */
  int32_t is_sorted;        /* 1: fwd holds (startDocId,endDocId) BE int pairs per dictId (SortedIndexReaderImpl) */
  int32_t dict_entry_bytes; /* 4/8 for numeric, numBytesPerValue (padded) for STRING */
  const uint8_t* fwd;
  int64_t fwd_len;
  const uint8_t* dict;
  int64_t dict_len;
  const uint8_t* inv;       /* bitmap inverted index file or NULL */
  int64_t inv_len;
} po_column_t;

typedef struct {
  int32_t num_docs;
  int32_t num_columns;
  const po_column_t* columns;
} po_segment_t;

/* ---- query (value space, like QueryContext) ---- */
enum { PO_AND = 0, PO_OR = 1, PO_NOT = 2, PO_EQ = 3, PO_NEQ = 4, PO_IN = 5, PO_NOT_IN = 6, PO_RANGE = 7,
       PO_DOCIDS = 8 /* BitmapBasedFilterOperator over po_query_t.doc_ids (star-tree traversal result) */ };

typedef struct {
  int64_t i;      /* INT / LONG literal */
  double d;       /* FLOAT / DOUBLE literal */
  const char* s;  /* STRING literal (NUL terminated) */
} po_literal_t;

/* Filter tree in POSTFIX order: children precede their parent; the last node is the root. */
typedef struct {
  int32_t type;          /* PO_AND .. PO_RANGE */
  int32_t column;        /* leaf: column index in the segment */
  int32_t num_children;  /* AND / OR (NOT has 1) */
  int32_t lower_inclusive, upper_inclusive, lower_unbounded, upper_unbounded; /* RANGE */
  int32_t num_values;    /* EQ/NEQ: 1; IN/NOT_IN: n; RANGE: 2 (lower, upper) */
  int32_t values_offset; /* first literal in po_query_t.literals */
} po_filter_node_t;

enum { PO_COUNT = 0, PO_SUM = 1, PO_MIN = 2, PO_MAX = 3, PO_AVG = 4, PO_DISTINCTCOUNT = 5 };

typedef struct {
  int32_t function; /* PO_COUNT .. */
  int32_t column;   /* -1 for COUNT(*) */
} po_agg_t;

typedef struct {
  int32_t num_filter_nodes; /* 0 = no filter */
  const po_filter_node_t* filter;
  const po_literal_t* literals;
  int32_t num_group_by;
  const int32_t* group_by_columns;
  int32_t num_aggs;
  const po_agg_t* aggs;
  int32_t num_groups_limit;                   /* InstancePlanMakerImplV2 default 100000 */
  int32_t max_initial_result_holder_capacity; /* default 10000 (= array based threshold) */
  int32_t and_scan_reordering;                /* query option AndScanReordering, default 0 */
  int64_t num_doc_ids;                        /* PO_DOCIDS leaf: sorted doc ids */
  const int32_t* doc_ids;
  /* FILTER (WHERE ...) per aggregation (QueryContext.getFilteredAggregationFunctions): aggregation a is filtered iff
   * agg_filter_count[a] > 0, its postfix tree is agg_filter_nodes[agg_filter_start[a] .. + agg_filter_count[a]) (literals
   * shared with the main filter); equal (start, count) = the same FILTER.  NULL arrays: no filtered aggregation. */
  const po_filter_node_t* agg_filter_nodes;
  const int32_t* agg_filter_start;
  const int32_t* agg_filter_count;
} po_query_t;

enum { PO_REGIME_NONE = 0, PO_REGIME_ARRAY = 1, PO_REGIME_INT_MAP = 2, PO_REGIME_LONG_MAP = 3, PO_REGIME_ARRAY_MAP = 4, PO_REGIME_NO_DICT = 5 };

typedef struct po_result po_result_t;

/* ---- formats ---- */
int32_t po_num_bits_per_value(int32_t max_value);
void po_bitset_write(uint8_t* buf, int64_t start_index, int32_t bits, int64_t n, const int32_t* values);
int32_t po_bitset_read(const uint8_t* buf, int64_t index, int32_t bits);
int32_t po_fixedbit_read(const uint8_t* buf, int64_t index, int32_t bits);
int32_t po_fixedbit_read_unchecked(const uint8_t* buf, int64_t index, int32_t bits);
void po_fixedbit_read32(const uint8_t* buf, int64_t index, int32_t bits, int32_t* out);
void po_fwd_read_dict_ids(const uint8_t* buf, int32_t num_docs, int32_t bits, const int32_t* doc_ids, int32_t length,
                          int32_t* out);

/* Roaring portable serialization.  Returns bytes written (or needed when out == NULL). */
int64_t po_roaring_serialize(const uint32_t* sorted_values, int64_t n, int32_t run_optimize, uint8_t* out,
                             int64_t cap);
/* Returns cardinality; writes up to cap values. -1 on malformed input. */
int64_t po_roaring_deserialize(const uint8_t* buf, int64_t len, uint32_t* out, int64_t cap);
/* Builds a bitmap inverted index file (BitmapInvertedIndexWriter layout) from per-doc dictIds. */
int64_t po_inverted_index_build(const int32_t* dict_ids, int32_t num_docs, int32_t cardinality, uint8_t* out,
                                int64_t cap);

/* ---- executor: one call == getOperator(query).nextBlock() on one segment ---- */
po_result_t* po_execute(const po_segment_t* segment, const po_query_t* query);
const char* po_result_error(const po_result_t* r); /* NULL when OK */
int32_t po_result_num_groups(const po_result_t* r); /* -1: aggregation only */
int32_t po_result_regime(const po_result_t* r);
int32_t po_result_groups_limit_reached(const po_result_t* r);
/* {numDocsScanned, numEntriesScannedInFilter, numEntriesScannedPostFilter, numTotalDocs} */
void po_result_stats(const po_result_t* r, int64_t out[4]);
/* [num_groups x num_group_by] dictIds, group-major */
void po_result_group_keys(const po_result_t* r, int32_t* out);
/* per aggregation: doubles (SUM/MIN/MAX value, AVG sum, COUNT as double), longs (COUNT, AVG count, DISTINCT size) */
void po_result_agg_double(const po_result_t* r, int32_t agg, double* out);
void po_result_agg_long(const po_result_t* r, int32_t agg, int64_t* out);
/* DISTINCTCOUNT: sorted dictIds of one group (group = 0 for aggregation only); returns count */
int64_t po_result_distinct(const po_result_t* r, int32_t agg, int32_t group, int32_t* out, int64_t cap);
int64_t po_result_raw_key_values(const po_result_t* r, int32_t group_by_column, double* out_d, int64_t* out_l, int64_t cap);
int64_t po_result_raw_distinct_values(const po_result_t* r, int32_t agg, double* out_d, int64_t* out_l, int64_t cap);
void po_result_free(po_result_t* r);

/* ---- star-tree (OffHeapStarTree / StarTreeFilterOperator.traverseStarTree) ---- */
typedef struct {
  int32_t dimension;     /* index into the tree's dimension split order */
  int32_t num_ids;
  const int32_t* ids;    /* matching dictIds of the (AND-ed) predicates on that dimension, sorted */
} po_star_predicate_t;
/* Header facts: out = {numDimensions, numNodes}; dimension names are copied NUL separated into names (cap bytes). */
int32_t po_startree_info(const uint8_t* tree, int64_t len, int32_t out[2], char* names, int32_t cap);
/* BFS traversal exactly as StarTreeFilterOperator: returns the number of matched star-tree docs (sorted ascending in
 * out_docs, up to cap), -1 when some predicate has no matching dictId (empty result), -2 on a malformed tree.
 * remaining_predicate_mask: bit d set = predicates on dimension d still have to be applied to the matched docs. */
int64_t po_startree_traverse(const uint8_t* tree, int64_t len, int32_t num_predicates,
                             const po_star_predicate_t* predicates, int32_t num_group_by, const int32_t* group_by_dims,
                             int32_t* out_docs, int64_t cap, uint32_t* remaining_predicate_mask);

/* Matching doc ids of the filter alone (tests). Returns count; writes up to cap. */
int64_t po_filter_doc_ids(const po_segment_t* segment, const po_query_t* query, int32_t* out, int64_t cap,
                          int64_t* entries_scanned_in_filter);

/* ---- benchmark table generator (CPU twin of pinot_b200/csrc/pb200_synth.cu; not part of the reference) ---- */
/* Forward index bytes of one synthetic dict-encoded INT column: dictId(doc) = mix64(seed + doc * golden) % cardinality,
 * written as an MSB-first big-endian bit stream of ceil(num_docs * bits / 8) bytes, by `threads` threads. */
void po_synth_fwd(uint64_t seed, int32_t cardinality, int64_t num_docs, int32_t bits, uint8_t* out, int32_t threads);
/* Dictionary { base + step * i }: cardinality big-endian INT values. */
void po_synth_dict(int32_t cardinality, int32_t base, int32_t step, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif
