// host_probe.cpp -- TEST shim: calls the product's host-side logic (inside libpinot_b200.so) without a GPU, so that the
// CPU test suite can compare it with the oracle:
//   * pb200h::StarTree::parse / traverse   (StarTreeFilterOperator's traversal as the product runs it)
//   * pb200h::matching_dict_ids             (PredicateEvaluator resolution: value-space predicate -> dictIds)
// Built by tests/test_host_logic_cpu.py with g++ against the in-tree library; nothing here ships.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../pinot_b200/csrc/host/host_internal.h"

extern "C" {

// preds: npreds entries (dimension, count, ids...) flattened as dims[i], offsets[i]..offsets[i+1] into ids[].
// Returns the number of [start, end) ranges written (pairs), -1 when a predicate matches nothing, -2 malformed tree.
int64_t probe_startree_traverse(const unsigned char* tree, uint64_t len, int32_t num_dims, int32_t npreds, const int32_t* dims,
                                const int32_t* offsets, const int32_t* ids, uint32_t group_by_mask, int32_t* out_pairs,
                                int64_t cap_pairs, uint32_t* remaining) {
  pb200h::StarTree t;
  if (!t.parse(tree, len) || (int)t.dim_names.size() != num_dims) return -2;
  std::vector<std::vector<int32_t>> store(num_dims);
  std::vector<const std::vector<int32_t>*> preds(num_dims, nullptr);
  for (int i = 0; i < npreds; i++) {
    if (dims[i] < 0 || dims[i] >= num_dims) return -2;
    store[dims[i]].assign(ids + offsets[i], ids + offsets[i + 1]);
    preds[dims[i]] = &store[dims[i]];
  }
  std::vector<std::pair<int32_t, int32_t>> docs;
  uint32_t rem = 0;
  if (!t.traverse(preds, group_by_mask, docs, rem)) return -1;
  if (remaining) *remaining = rem;
  int64_t n = 0;
  for (auto& r : docs) {
    if (n < cap_pairs) { out_pairs[2 * n] = r.first; out_pairs[2 * n + 1] = r.second; }
    n++;
  }
  return n;
}

// One predicate against one dictionary column described like pb200h_segment_create's input.
int64_t probe_matching_dict_ids(const pb200h_column* col, const pb200h_filter_node* node, const pb200h_literal* literals,
                                int32_t* out, int64_t cap) {
  pb200h::HostColumn h;
  h.name = col->name ? col->name : "";
  h.data_type = col->data_type; h.has_dictionary = col->has_dictionary; h.bits = col->bits_per_value;
  h.cardinality = col->cardinality; h.is_sorted = col->is_sorted; h.entry_bytes = col->dict_entry_bytes;
  if (col->dict && col->dict_bytes) h.dict.assign((const unsigned char*)col->dict, (const unsigned char*)col->dict + col->dict_bytes);
  std::vector<int32_t> ids = pb200h::matching_dict_ids(h, *node, literals);
  for (size_t i = 0; i < ids.size() && (int64_t)i < cap; i++) out[i] = ids[i];
  return (int64_t)ids.size();
}

// FilterOperatorUtils.getLeafFilterOperator as the product runs it: one predicate on one column of a (host-only) segment ->
// the device leaf.  out4 = {op, column, lo, hi}; ids (<= cap) receives the node's id / doc-range payload; returns num_ids
// or a negative pb200 status.
int64_t probe_leaf_to_device(const pb200h_column* col, int32_t num_docs, const pb200h_filter_node* node,
                             const pb200h_literal* literals, int32_t* out4, int32_t* ids, int64_t cap) {
  pb200h_segment seg;
  seg.num_docs = num_docs;
  pb200h::HostColumn h;
  h.name = col->name ? col->name : "";
  h.data_type = col->data_type; h.has_dictionary = col->has_dictionary; h.bits = col->bits_per_value;
  h.cardinality = col->cardinality; h.is_sorted = col->is_sorted; h.entry_bytes = col->dict_entry_bytes;
  h.has_inverted = col->inv != nullptr && col->inv_bytes > 0 && !col->is_sorted;
  if (col->dict && col->dict_bytes) h.dict.assign((const unsigned char*)col->dict, (const unsigned char*)col->dict + col->dict_bytes);
  if (col->is_sorted && col->fwd) h.sorted_idx.assign((const unsigned char*)col->fwd, (const unsigned char*)col->fwd + col->fwd_bytes);
  seg.cols.push_back(std::move(h));
  pb200h::SegmentFilterStore store;
  pb200_filter_node d;
  int rc = pb200h::leaf_to_device(seg, 0, *node, literals, store, d);
  if (rc) return rc;
  out4[0] = d.op; out4[1] = d.column; out4[2] = d.lo; out4[3] = d.hi;
  for (int i = 0; i < d.num_ids && i < cap; i++) ids[i] = d.ids[i];
  return d.num_ids;
}

// pb200h_explain over a HOST-ONLY segment (no device registration): the filter tree after predicate resolution and
// constant folding (FilterPlanNode / AndFilterOperator / OrFilterOperator always-true / always-false rules) plus the
// operator the plan maker would choose.  Returns the text length or a negative status.
int32_t probe_explain(const pb200h_column* cols, int32_t ncols, int32_t num_docs, const pb200h_query* q, char* out, int32_t cap) {
  pb200h_segment seg;
  seg.num_docs = num_docs;
  for (int i = 0; i < ncols; i++) {
    const pb200h_column& c = cols[i];
    pb200h::HostColumn h;
    h.name = c.name ? c.name : "";
    h.data_type = c.data_type; h.has_dictionary = c.has_dictionary; h.bits = c.bits_per_value;
    h.cardinality = c.cardinality; h.is_sorted = c.is_sorted; h.entry_bytes = c.dict_entry_bytes;
    h.has_inverted = c.inv != nullptr && c.inv_bytes > 0 && !c.is_sorted;
    if (c.dict && c.dict_bytes) h.dict.assign((const unsigned char*)c.dict, (const unsigned char*)c.dict + c.dict_bytes);
    if (c.is_sorted && c.fwd) h.sorted_idx.assign((const unsigned char*)c.fwd, (const unsigned char*)c.fwd + c.fwd_bytes);
    seg.cols.push_back(std::move(h));
  }
  return pb200h_explain(nullptr, q, &seg, out, cap);
}

// raw_forward.cpp: every chunk of a fixed-width raw SV forward index decoded into big-endian values (out: num_docs * width B)
int32_t probe_decode_fixed_byte_forward(const unsigned char* file, uint64_t len, int32_t width, int64_t num_docs, unsigned char* out) {
  std::vector<unsigned char> v;
  int rc = pb200h::decode_fixed_byte_forward(file, len, width, num_docs, v);
  if (rc) return rc;
  memcpy(out, v.data(), v.size());
  return 0;
}

// raw_forward.cpp: dictionary synthesised from raw values.  Returns 1 (built), 0 (more than max_cardinality distinct values).
int32_t probe_synthesize_dictionary(const unsigned char* values_be, int32_t data_type, int64_t num_docs, int32_t max_cardinality,
                                    unsigned char* dict_out, int64_t dict_cap, unsigned char* fwd_out, int64_t fwd_cap,
                                    int32_t* cardinality, int32_t* bits) {
  const int w = pb200h::raw_value_width(data_type);
  std::vector<unsigned char> values(values_be, values_be + (size_t)num_docs * w), dict, fwd;
  int card = 0, nb = 0;
  if (!pb200h::synthesize_dictionary(values, data_type, num_docs, max_cardinality, dict, fwd, &card, &nb)) return 0;
  if ((int64_t)dict.size() > dict_cap || (int64_t)fwd.size() > fwd_cap) return -1;
  memcpy(dict_out, dict.data(), dict.size());
  memcpy(fwd_out, fwd.data(), fwd.size());
  *cardinality = card; *bits = nb;
  return 1;
}

}  // extern "C"
