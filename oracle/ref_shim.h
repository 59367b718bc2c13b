/* Forced-include shim (g++ -include) used ONLY to compile the unmodified
 * reference sources in place with GCC 13 / libstdc++ 13.
 *
 * libstdc++ >= 11 compares allocators when list/map nodes are spliced or
 * moved; the reference's Pool_alloc (src/memoryPool/pool_alloc.h:12-91)
 * predates that requirement and defines no operator==/!=.  Declaring the
 * class template ahead of its definition lets us supply the two operators
 * without touching the reference tree.  (The second incompatibility, the
 * map allocator's value_type at src/branch_and_bound.h:44, is only a
 * static_assert under __STRICT_ANSI__, so the recipe builds with
 * -std=gnu++11 instead of -std=c++11.)
 */
#ifndef XXM_REF_SHIM_H
#define XXM_REF_SHIM_H
#ifdef __cplusplus
template <class T> class Pool_alloc;
template <class T, class U>
inline bool operator==(const Pool_alloc<T>&, const Pool_alloc<U>&) { return true; }
template <class T, class U>
inline bool operator!=(const Pool_alloc<T>&, const Pool_alloc<U>&) { return false; }
#endif
#endif
