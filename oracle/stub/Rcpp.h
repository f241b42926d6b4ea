/* Stand-in for the parts of <Rcpp.h> that /root/reference/src/autoFRK.cpp uses -- TEST INFRASTRUCTURE.
 *
 *   Rcpp::stop(fmt, ...)        -> throws std::runtime_error carrying the formatted message (Rcpp formats with
 *                                  tinyformat, whose %d / %s output equals printf's for the int arguments used);
 *   Rcpp::List::create(Named(..) = matrix, ...) -> an ordered name -> matrix list the C wrapper reads back.
 * No R session exists behind it; oracle/ref_wrap.cpp turns the exception into a status + message.
 */
#ifndef FRK_STUB_RCPP_H
#define FRK_STUB_RCPP_H
#include <cstdio>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "frk_stub_eigen.h"

namespace Rcpp {

struct exception : public std::runtime_error {
    explicit exception(const std::string &m) : std::runtime_error(m) {}
};

inline void stop(const char *msg) { throw exception(msg); }
template <typename... Args> inline void stop(const char *fmt, Args... args) {
    char buf[512];
    std::snprintf(buf, sizeof buf, fmt, args...);
    throw exception(buf);
}

struct NamedValue {
    std::string name;
    Eigen::MatrixXd value;
};
class Named {
    std::string name_;

  public:
    explicit Named(const char *n) : name_(n) {}
    NamedValue operator=(const Eigen::MatrixXd &m) const { return NamedValue{name_, m}; }
};

class List {
    std::vector<NamedValue> items_;

  public:
    template <typename... Items> static List create(Items &&...it) {
        List l;
        (l.items_.push_back(std::forward<Items>(it)), ...);
        return l;
    }
    size_t size() const { return items_.size(); }
    const NamedValue &at(size_t i) const { return items_.at(i); }
    const Eigen::MatrixXd *find(const std::string &name) const {
        for (const NamedValue &nv : items_)
            if (nv.name == name) return &nv.value;
        return nullptr;
    }
};

}  // namespace Rcpp
#endif
