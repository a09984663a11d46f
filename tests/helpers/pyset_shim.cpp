// Host build of csrc/pyset.h for the CPU test that checks the set-order emulation
// against the real CPython set (tests/test_pyset.py).
#include "../../rag_search_engine_b200/csrc/pyset.h"
#include <vector>
extern "C" int pyset_shim_union(const int64_t* a, int na, const int64_t* b, int nb, int64_t* out) {
  int ca = rse::pyset_capacity_for(na), cb = rse::pyset_capacity_for(nb), cr = rse::pyset_capacity_for(na + nb);
  std::vector<int64_t> ta(ca), tb(cb), tr(cr), sc(cr);
  return rse::pyset_union_order(a, na, b, nb, ta.data(), ca, tb.data(), cb, tr.data(), cr, sc.data(), cr, out);
}
