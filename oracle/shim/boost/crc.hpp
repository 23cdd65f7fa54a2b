/* oracle/shim/boost/crc.hpp -- TEST INFRASTRUCTURE ONLY.
 * boost::crc_32_type as the reference uses it for checkpoint files (src/miscellaneous.cc:396,444):
 * the standard reflected CRC-32 (polynomial 0xEDB88320, init and final xor 0xFFFFFFFF). */
#ifndef QB_ORACLE_SHIM_BOOST_CRC_HPP
#define QB_ORACLE_SHIM_BOOST_CRC_HPP
#include <cstddef>
#include <cstdint>
namespace boost {
class crc_32_type {
public:
    crc_32_type() : rem_(0xFFFFFFFFu) {}
    void process_bytes(const void *buf, std::size_t len) {
        static const Table tab;
        const unsigned char *p = static_cast<const unsigned char *>(buf);
        for (std::size_t i = 0; i < len; i++) rem_ = tab.t[(rem_ ^ p[i]) & 0xFFu] ^ (rem_ >> 8);
    }
    std::uint32_t checksum() const { return rem_ ^ 0xFFFFFFFFu; }
private:
    struct Table { std::uint32_t t[256]; Table() { for (std::uint32_t i = 0; i < 256; i++) { std::uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1); t[i] = c; } } };
    std::uint32_t rem_;
};
}
#endif
