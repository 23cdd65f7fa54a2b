/* oracle/shim/boost/asio.hpp -- TEST INFRASTRUCTURE ONLY: the one Boost.Asio call the reference makes
 * (ip::host_name(), src/miscellaneous.cc:51). */
#ifndef QB_ORACLE_SHIM_BOOST_ASIO_HPP
#define QB_ORACLE_SHIM_BOOST_ASIO_HPP
#include <string>
#include <unistd.h>
namespace boost { namespace asio { namespace ip {
inline std::string host_name() { char buf[256] = {0}; if (gethostname(buf, sizeof(buf) - 1) != 0) return "unknown"; return buf; }
}}}
#endif
