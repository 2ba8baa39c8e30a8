// Stand-in for the OptiX 6.5 SDK header <optixu/optixu_matrix_namespace.h>.
// TEST INFRASTRUCTURE ONLY; our own code, not OptiX. Row-major M x N float matrix with the
// handful of members the reference's host-compilable headers use.
#ifndef BPT_ORACLE_OPTIXU_MATRIX_NAMESPACE_H
#define BPT_ORACLE_OPTIXU_MATRIX_NAMESPACE_H

#include <optixu/optixu_math_namespace.h>

namespace optix {

template <int N> struct VectorDim {};
template <> struct VectorDim<2> { typedef float2 VectorType; };
template <> struct VectorDim<3> { typedef float3 VectorType; };
template <> struct VectorDim<4> { typedef float4 VectorType; };

template <unsigned int M, unsigned int N>
class Matrix {
public:
    typedef typename VectorDim<N>::VectorType floatN; // A row
    typedef typename VectorDim<M>::VectorType floatM; // A column

    Matrix() = default;
    explicit Matrix(const float data[M * N]) { for (unsigned int i = 0; i < M * N; ++i) m_data[i] = data[i]; }

    float operator[](unsigned int i) const { return m_data[i]; }
    float& operator[](unsigned int i) { return m_data[i]; }
    float* getData() { return m_data; }
    const float* getData() const { return m_data; }

    floatN getRow(unsigned int m) const {
        floatN r; float* v = reinterpret_cast<float*>(&r);
        for (unsigned int i = 0; i < N; ++i) v[i] = m_data[m * N + i];
        return r;
    }
    floatM getCol(unsigned int n) const {
        floatM c; float* v = reinterpret_cast<float*>(&c);
        for (unsigned int i = 0; i < M; ++i) v[i] = m_data[i * N + n];
        return c;
    }
    void setRow(unsigned int m, const floatN& r) {
        const float* v = reinterpret_cast<const float*>(&r);
        for (unsigned int i = 0; i < N; ++i) m_data[m * N + i] = v[i];
    }
    void setCol(unsigned int n, const floatM& c) {
        const float* v = reinterpret_cast<const float*>(&c);
        for (unsigned int i = 0; i < M; ++i) m_data[i * N + n] = v[i];
    }

    Matrix<N, M> transpose() const {
        Matrix<N, M> t;
        for (unsigned int r = 0; r < M; ++r)
            for (unsigned int c = 0; c < N; ++c)
                t[c * M + r] = m_data[r * N + c];
        return t;
    }

    static Matrix<M, N> identity() {
        Matrix<M, N> id;
        for (unsigned int r = 0; r < M; ++r)
            for (unsigned int c = 0; c < N; ++c)
                id[r * N + c] = r == c ? 1.0f : 0.0f;
        return id;
    }

private:
    float m_data[M * N];
};

typedef Matrix<2, 2> Matrix2x2;
typedef Matrix<3, 3> Matrix3x3;
typedef Matrix<4, 4> Matrix4x4;

OPTIXU_INLINE float2 operator*(const Matrix2x2& m, const float2& v) {
    return ::make_float2(m[0] * v.x + m[1] * v.y,
                         m[2] * v.x + m[3] * v.y);
}
OPTIXU_INLINE float3 operator*(const Matrix3x3& m, const float3& v) {
    return ::make_float3(m[0] * v.x + m[1] * v.y + m[2] * v.z,
                         m[3] * v.x + m[4] * v.y + m[5] * v.z,
                         m[6] * v.x + m[7] * v.y + m[8] * v.z);
}
OPTIXU_INLINE float4 operator*(const Matrix4x4& m, const float4& v) {
    return ::make_float4(m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * v.w,
                         m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7] * v.w,
                         m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11] * v.w,
                         m[12] * v.x + m[13] * v.y + m[14] * v.z + m[15] * v.w);
}

} // namespace optix

#endif // BPT_ORACLE_OPTIXU_MATRIX_NAMESPACE_H
