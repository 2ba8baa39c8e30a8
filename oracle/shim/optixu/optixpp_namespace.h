// Stand-in for <optixu/optixpp_namespace.h>. TEST INFRASTRUCTURE ONLY; our own code.
// The reference's host-compiled headers only need optix::Exception to exist.
#ifndef BPT_ORACLE_OPTIXPP_NAMESPACE_H
#define BPT_ORACLE_OPTIXPP_NAMESPACE_H

#include <string>

namespace optix {
class Exception {
public:
    explicit Exception(const std::string& message = "", int code = 0) : m_message(message), m_code(code) {}
    const std::string& getErrorString() const { return m_message; }
    int getErrorCode() const { return m_code; }
private:
    std::string m_message;
    int m_code;
};
} // namespace optix

#endif // BPT_ORACLE_OPTIXPP_NAMESPACE_H
