/* oracle/shim/boost/version.hpp -- TEST INFRASTRUCTURE ONLY (src/miscellaneous.cc:71,107). */
#ifndef QB_ORACLE_SHIM_BOOST_VERSION_HPP
#define QB_ORACLE_SHIM_BOOST_VERSION_HPP
#define BOOST_LIB_VERSION "shim"
#define BOOST_PLATFORM "linux (shim)"
#endif
