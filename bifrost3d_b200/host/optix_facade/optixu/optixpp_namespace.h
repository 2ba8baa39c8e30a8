// Look-alike facade for the two OptiX types that cross OptiXRenderer::Renderer's public API
// (extensions/OptiXRenderer/OptiXRenderer/Renderer.h:22-28,67,71): optix::Buffer (the half4 render target the caller
// maps, tests/OptiXRendererTests/RendererTest.h:115-128) and optix::Context (only used to create that buffer).
// This is NOT OptiX and calls no OptiX: a Buffer is a ref-counted CUDA device allocation.
#ifndef BPT_OPTIX_FACADE_OPTIXPP_NAMESPACE_H
#define BPT_OPTIX_FACADE_OPTIXPP_NAMESPACE_H

#include <cstddef>
#include <string>
#include <vector>

enum RTbuffertype { RT_BUFFER_INPUT = 1, RT_BUFFER_OUTPUT = 2, RT_BUFFER_INPUT_OUTPUT = 3 };
enum RTformat { RT_FORMAT_UNKNOWN = 0x100, RT_FORMAT_FLOAT4 = 0x104, RT_FORMAT_HALF4 = 0x12d };
typedef size_t RTsize;

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

// Intrusive ref-counted handle, like optix::Handle<T>.
template <class T>
class Handle {
public:
    Handle() : m_ptr(nullptr) {}
    Handle(T* ptr) : m_ptr(ptr) { if (m_ptr) m_ptr->add_reference(); }
    Handle(const Handle& other) : m_ptr(other.m_ptr) { if (m_ptr) m_ptr->add_reference(); }
    ~Handle() { if (m_ptr && m_ptr->remove_reference() == 0) delete m_ptr; }
    Handle& operator=(const Handle& other) {
        if (other.m_ptr) other.m_ptr->add_reference();
        if (m_ptr && m_ptr->remove_reference() == 0) delete m_ptr;
        m_ptr = other.m_ptr;
        return *this;
    }
    T* operator->() const { return m_ptr; }
    T* get() const { return m_ptr; }
    operator bool() const { return m_ptr != nullptr; }
private:
    T* m_ptr;
};

class RefCounted {
public:
    void add_reference() { ++m_references; }
    int remove_reference() { return --m_references; }
    virtual ~RefCounted() {}
private:
    int m_references = 0;
};

// Device buffer of width x height elements. map() copies it to host memory and returns that copy.
class BufferObj : public RefCounted {
public:
    BufferObj(RTformat format, RTsize width, RTsize height);
    ~BufferObj() override;
    void* map();
    void unmap() {}
    void getSize(RTsize& width, RTsize& height) const { width = m_width; height = m_height; }
    RTformat getFormat() const { return m_format; }
    RTsize getElementSize() const { return m_element_size; }
    void setDevicePointer(int device, void* pointer); // caller-owned device memory (DX11OptiXAdaptor/Adaptor.cpp:175)
    void* getDevicePointer(int device) const { return m_device; }
    int getId() const { return m_id; }
    void destroy() {}
private:
    RTformat m_format;
    RTsize m_width, m_height, m_element_size;
    void* m_device;
    bool m_owns_device;
    std::vector<unsigned char> m_host;
    int m_id;
};
typedef Handle<BufferObj> Buffer;

class ContextObj : public RefCounted {
public:
    Buffer createBuffer(unsigned int type, RTformat format, RTsize width, RTsize height) { return Buffer(new BufferObj(format, width, height)); }
};
typedef Handle<ContextObj> Context;

} // namespace optix

#endif // BPT_OPTIX_FACADE_OPTIXPP_NAMESPACE_H
