/*
    enoki/autodiff_b200.h -- members of enoki::Tape<CUDAArray<float|double>> for the enoki_b200 backend.

    The reference's DiffArray<T> (include/enoki/autodiff.h:126-1412, used UNMODIFIED) talks to its tape
    exclusively through the private members declared in autodiff.h:23-124, which the reference defines
    in src/autodiff/autodiff.cpp and explicitly instantiates in libenoki-autodiff.so.  This header
    supplies those members for CUDAArray value types as explicit specialisations that forward to the
    C ABI tape runtime of libenoki_b200.so (ek_tape_*, include/enoki_b200.h), so
    `DiffArray<CUDAArray<float>>` links against libenoki_b200.so instead of libenoki-autodiff.so.

    Included automatically by this repository's <enoki/cuda.h> when <enoki/autodiff.h> was included first.
*/
#pragma once

#include <enoki/autodiff.h>
#include <unordered_map>

NAMESPACE_BEGIN(enoki)

namespace detail {
    template <typename T> struct ek_tape_type;
    template <> struct ek_tape_type<CUDAArray<float>>  { static constexpr ek_type value = EK_FLOAT32; };
    template <> struct ek_tape_type<CUDAArray<double>> { static constexpr ek_type value = EK_FLOAT64; };
    /// gradient() hands out `const Type &`: keep the most recently requested gradients alive here
    template <typename T> std::unordered_map<uint32_t, T> &ek_grad_cache() {
        static std::unordered_map<uint32_t, T> cache;
        return cache;
    }
}

#define ENOKI_B200_TAPE(Type)                                                                          \
    template <> inline Tape<Type>::Tape() : d(nullptr) { }                                             \
    template <> inline Tape<Type>::~Tape() { }                                                         \
    template <> inline Tape<Type> *Tape<Type>::get() {                                                 \
        static Tape<Type> *instance = new Tape<Type>();                                                \
        return instance;                                                                               \
    }                                                                                                  \
    template <> inline uint32_t Tape<Type>::append(const char *label, size_t size, Index i1, const Type &w1) { \
        uint32_t in[1] = { i1 }, w[1] = { w1.index_() };                                               \
        return ek_tape_append(detail::ek_tape_type<Type>::value, label, size, 1, in, w);               \
    }                                                                                                  \
    template <> inline uint32_t Tape<Type>::append(const char *label, size_t size, Index i1, Index i2, \
                                                  const Type &w1, const Type &w2) {                    \
        uint32_t in[2] = { i1, i2 }, w[2] = { w1.index_(), w2.index_() };                              \
        return ek_tape_append(detail::ek_tape_type<Type>::value, label, size, 2, in, w);               \
    }                                                                                                  \
    template <> inline uint32_t Tape<Type>::append(const char *label, size_t size, Index i1, Index i2, Index i3, \
                                                  const Type &w1, const Type &w2, const Type &w3) {    \
        uint32_t in[3] = { i1, i2, i3 }, w[3] = { w1.index_(), w2.index_(), w3.index_() };             \
        return ek_tape_append(detail::ek_tape_type<Type>::value, label, size, 3, in, w);               \
    }                                                                                                  \
    template <> inline uint32_t Tape<Type>::append_psum(Index i) { return ek_tape_append_psum(detail::ek_tape_type<Type>::value, i); } \
    template <> inline uint32_t Tape<Type>::append_reverse(Index i) { return ek_tape_append_reverse(detail::ek_tape_type<Type>::value, i); } \
    template <> inline uint32_t Tape<Type>::append_gather(const Int64 &offset, const Mask &mask) {     \
        return ek_tape_append_gather(detail::ek_tape_type<Type>::value, offset.index_(), mask.index_()); \
    }                                                                                                  \
    template <> inline void Tape<Type>::append_scatter(Index index, const Int64 &offset, const Mask &mask, bool scatter_add) { \
        if (ek_tape_append_scatter(detail::ek_tape_type<Type>::value, index, offset.index_(), mask.index_(), scatter_add) != 0) \
            detail::ek_throw();                                                                        \
    }                                                                                                  \
    template <> inline uint32_t Tape<Type>::append_node(size_t size, const char *label) {              \
        return ek_tape_append_node(detail::ek_tape_type<Type>::value, size, label);                    \
    }                                                                                                  \
    template <> inline uint32_t Tape<Type>::append_leaf(size_t size) { return ek_tape_append_leaf(detail::ek_tape_type<Type>::value, size); } \
    template <> inline void Tape<Type>::append_edge(Index src, Index dst, const Type &weight) {        \
        if (ek_tape_append_edge(detail::ek_tape_type<Type>::value, src, dst, weight.index_()) != 0) detail::ek_throw(); \
    }                                                                                                  \
    template <> inline void Tape<Type>::dec_ref_ext(Index index) { ek_tape_dec_ref_ext(detail::ek_tape_type<Type>::value, index); } \
    template <> inline void Tape<Type>::inc_ref_ext(Index index) { ek_tape_inc_ref_ext(detail::ek_tape_type<Type>::value, index); } \
    template <> inline void Tape<Type>::set_scatter_gather_operand(Index *index, size_t size, bool permute) { \
        if (ek_tape_set_scatter_gather_operand(detail::ek_tape_type<Type>::value, index, size, permute) != 0) detail::ek_throw(); \
    }                                                                                                  \
    template <> inline void Tape<Type>::push_prefix(const char *p) { ek_tape_push_prefix(detail::ek_tape_type<Type>::value, p); } \
    template <> inline void Tape<Type>::pop_prefix() { if (ek_tape_pop_prefix(detail::ek_tape_type<Type>::value) != 0) detail::ek_throw(); } \
    template <> inline void Tape<Type>::backward(bool free_graph) {                                    \
        detail::ek_grad_cache<Type>().clear();                                                         \
        if (ek_tape_backward_static(detail::ek_tape_type<Type>::value, free_graph) != 0) detail::ek_throw(); \
    }                                                                                                  \
    template <> inline void Tape<Type>::forward(bool free_graph) {                                     \
        detail::ek_grad_cache<Type>().clear();                                                         \
        if (ek_tape_forward_static(detail::ek_tape_type<Type>::value, free_graph) != 0) detail::ek_throw(); \
    }                                                                                                  \
    template <> inline void Tape<Type>::backward(Index index, bool free_graph) {                       \
        detail::ek_grad_cache<Type>().clear();                                                         \
        if (ek_tape_backward(detail::ek_tape_type<Type>::value, index, free_graph) != 0) detail::ek_throw(); \
    }                                                                                                  \
    template <> inline void Tape<Type>::forward(Index index, bool free_graph) {                        \
        detail::ek_grad_cache<Type>().clear();                                                         \
        if (ek_tape_forward(detail::ek_tape_type<Type>::value, index, free_graph) != 0) detail::ek_throw(); \
    }                                                                                                  \
    template <> inline void Tape<Type>::set_gradient(Index index, const Type &value, bool backward) {  \
        if (ek_tape_set_gradient(detail::ek_tape_type<Type>::value, index, value.index_(), backward) != 0) detail::ek_throw(); \
    }                                                                                                  \
    template <> inline void Tape<Type>::set_label(Index index, const char *name) {                     \
        ek_tape_set_label(detail::ek_tape_type<Type>::value, index, name);                             \
    }                                                                                                  \
    template <> inline const Type &Tape<Type>::gradient(Index index) {                                 \
        uint32_t h = ek_tape_gradient(detail::ek_tape_type<Type>::value, index);                       \
        if (h == 0 && *ek_last_error()) detail::ek_throw();                                            \
        Type &slot = detail::ek_grad_cache<Type>()[index];                                             \
        if (h) { ek_inc_ref_ext(h); slot = Type::from_index_(h); } else slot = Type();                 \
        return slot;                                                                                   \
    }                                                                                                  \
    template <> inline std::string Tape<Type>::graphviz(const std::vector<Index> &indices) {           \
        char *s = ek_tape_graphviz(detail::ek_tape_type<Type>::value, indices.size(), indices.data()); \
        std::string r(s ? s : ""); free(s); return r;                                                  \
    }                                                                                                  \
    template <> inline void Tape<Type>::set_log_level(uint32_t l) { ek_tape_set_log_level(detail::ek_tape_type<Type>::value, l); } \
    template <> inline uint32_t Tape<Type>::log_level() const { return 0; }                            \
    template <> inline void Tape<Type>::set_graph_simplification(bool v) { ek_tape_set_graph_simplification(detail::ek_tape_type<Type>::value, v); } \
    template <> inline void Tape<Type>::simplify_graph() { ek_tape_simplify(detail::ek_tape_type<Type>::value); } \
    template <> inline std::string Tape<Type>::whos() const {                                          \
        char *s = ek_tape_whos(detail::ek_tape_type<Type>::value); std::string r(s ? s : ""); free(s); return r; \
    }

ENOKI_B200_TAPE(CUDAArray<float>)
ENOKI_B200_TAPE(CUDAArray<double>)

#undef ENOKI_B200_TAPE

NAMESPACE_END(enoki)
