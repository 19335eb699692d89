/* Stand-in for <boost/program_options.hpp>: the reference's logger.h names variables_map in one
 * signature (Logger::parseLogArgs); the oracle driver never calls it. */
#ifndef MB_ORACLE_STUB_BOOST_PO_H
#define MB_ORACLE_STUB_BOOST_PO_H
#include <map>
#include <string>
#include <vector>
#include <stdexcept>
namespace boost { namespace program_options {
struct variable_value {
  template<class T> const T& as() const { throw std::runtime_error ("oracle stub: boost::program_options is not available"); }
};
struct variables_map : std::map<std::string, variable_value> { };
} }
#endif
