/* Test infrastructure only (oracle/): tiny polymorphic-value stand-in covering
 * exactly the pmt:: calls the reference makes (preamble_impl.cc:52-53,100-137,
 * 227-232; slicer_impl.cc:112-114,184-192). */
#ifndef ORACLE_SHIM_PMT_H
#define ORACLE_SHIM_PMT_H
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
namespace pmt {
struct pmt_node {
    enum kind_t { SYMBOL, U64, F64, TUPLE } kind;
    std::string sym;
    uint64_t u = 0;
    double d = 0.0;
    std::vector<std::shared_ptr<pmt_node>> elems;
};
typedef std::shared_ptr<pmt_node> pmt_t;
inline pmt_t string_to_symbol(const std::string& s) {
    pmt_t p = std::make_shared<pmt_node>(); p->kind = pmt_node::SYMBOL; p->sym = s; return p;
}
inline bool is_symbol(const pmt_t& p) { return p && p->kind == pmt_node::SYMBOL; }
inline std::string symbol_to_string(const pmt_t& p) { return p->sym; }
inline pmt_t from_uint64(uint64_t v) {
    pmt_t p = std::make_shared<pmt_node>(); p->kind = pmt_node::U64; p->u = v; return p;
}
inline pmt_t from_double(double v) {
    pmt_t p = std::make_shared<pmt_node>(); p->kind = pmt_node::F64; p->d = v; return p;
}
inline uint64_t to_uint64(const pmt_t& p) { return p->u; }
inline double to_double(const pmt_t& p) { return p->d; }
inline pmt_t make_tuple(const pmt_t& a, const pmt_t& b) {
    pmt_t p = std::make_shared<pmt_node>(); p->kind = pmt_node::TUPLE;
    p->elems.push_back(a); p->elems.push_back(b); return p;
}
inline pmt_t tuple_ref(const pmt_t& t, size_t k) { return t->elems.at(k); }
}  // namespace pmt
#endif
