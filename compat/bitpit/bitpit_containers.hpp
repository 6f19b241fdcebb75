// bitpit_containers.hpp -- minimal stand-in for bitpit's pierced containers (see README.md).
// Ids are dense here (id == raw position, no holes): that is what a freshly built, never adapted
// VolOctree has, and it is all minimmerflow relies on.
#ifndef MMF_COMPAT_BITPIT_CONTAINERS_HPP
#define MMF_COMPAT_BITPIT_CONTAINERS_HPP

#include "bitpit_common.hpp"

#include <type_traits>

namespace bitpit {

template <typename id_t = long>
class PiercedKernel {
public:
    virtual ~PiercedKernel() = default;
    std::size_t size() const { return m_size; }
    std::size_t rawSize() const { return m_size; }
    bool contains(id_t id) const { return id >= 0 && (std::size_t) id < m_size; }
    std::size_t getRawIndex(id_t id) const { return (std::size_t) id; }

protected:
    std::size_t m_size = 0;
};

template <typename value_t, typename id_t, bool is_const>
class PiercedIterator {
public:
    typedef typename std::conditional<is_const, const value_t, value_t>::type elem_t;
    PiercedIterator() = default;
    PiercedIterator(elem_t *base, std::size_t pos) : m_base(base), m_pos(pos) {}
    id_t getId() const { return (id_t) m_pos; }
    std::size_t getRawIndex() const { return m_pos; }
    elem_t &operator*() const { return m_base[m_pos]; }
    elem_t *operator->() const { return m_base + m_pos; }
    PiercedIterator &operator++() { ++m_pos; return *this; }
    PiercedIterator operator++(int) { PiercedIterator t(*this); ++m_pos; return t; }
    bool operator==(const PiercedIterator &o) const { return m_pos == o.m_pos; }
    bool operator!=(const PiercedIterator &o) const { return m_pos != o.m_pos; }

private:
    elem_t *m_base = nullptr;
    std::size_t m_pos = 0;
};

template <typename value_t, typename id_t = long>
class PiercedVector : public PiercedKernel<id_t> {
public:
    typedef PiercedIterator<value_t, id_t, false> iterator;
    typedef PiercedIterator<value_t, id_t, true> const_iterator;

    value_t &rawAt(std::size_t pos) { return m_items[pos]; }
    const value_t &rawAt(std::size_t pos) const { return m_items[pos]; }
    value_t &at(id_t id) { return m_items.at((std::size_t) id); }
    const value_t &at(id_t id) const { return m_items.at((std::size_t) id); }
    value_t &operator[](id_t id) { return m_items[(std::size_t) id]; }
    const value_t &operator[](id_t id) const { return m_items[(std::size_t) id]; }

    iterator begin() { return iterator(m_items.data(), 0); }
    iterator end() { return iterator(m_items.data(), m_items.size()); }
    const_iterator begin() const { return const_iterator(m_items.data(), 0); }
    const_iterator end() const { return const_iterator(m_items.data(), m_items.size()); }
    const_iterator cbegin() const { return begin(); }
    const_iterator cend() const { return end(); }
    const_iterator find(id_t id) const { return const_iterator(m_items.data(), (std::size_t) id); }
    const_iterator rawFind(std::size_t pos) const { return const_iterator(m_items.data(), pos); }
    iterator rawFind(std::size_t pos) { return iterator(m_items.data(), pos); }

    // construction side (used by the mesh only)
    void reserve(std::size_t n) { m_items.reserve(n); }
    void clear() { m_items.clear(); this->m_size = 0; }
    template <typename... Args>
    value_t &emplaceBack(Args &&...args)
    {
        m_items.emplace_back(std::forward<Args>(args)...);
        this->m_size = m_items.size();
        return m_items.back();
    }

private:
    std::vector<value_t> m_items;
};

// AoS field storage attached to a kernel: value (raw position p, field k) lives at [p*nFields + k].
template <typename value_t, typename id_t = long>
class PiercedStorage {
    // bool is stored one byte per flag (bitpit packs it; callers only use rawAt/rawSet/at on it)
    typedef typename std::conditional<std::is_same<value_t, bool>::value, unsigned char, value_t>::type store_t;

public:
    PiercedStorage() = default;
    PiercedStorage(std::size_t nFields, const PiercedKernel<id_t> *kernel) : m_nFields(nFields) { setStaticKernel(kernel); }

    void setStaticKernel(const PiercedKernel<id_t> *kernel)
    {
        m_kernel = kernel;
        m_data.assign(m_nFields * (kernel ? kernel->rawSize() : 0), store_t());
    }
    const PiercedKernel<id_t> *getKernel() const { return m_kernel; }
    std::size_t getFieldCount() const { return m_nFields; }
    std::size_t rawSize() const { return m_nFields ? m_data.size() / m_nFields : 0; }

    store_t &rawAt(std::size_t pos, std::size_t k = 0) { return m_data[pos * m_nFields + k]; }
    const store_t &rawAt(std::size_t pos, std::size_t k = 0) const { return m_data[pos * m_nFields + k]; }
    void rawSet(std::size_t pos, const value_t &value) { m_data[pos * m_nFields] = (store_t) value; }
    void rawSet(std::size_t pos, std::size_t k, const value_t &value) { m_data[pos * m_nFields + k] = (store_t) value; }
    store_t *rawData(std::size_t pos, std::size_t offset = 0) { return m_data.data() + pos * m_nFields + offset; }
    const store_t *rawData(std::size_t pos, std::size_t offset = 0) const { return m_data.data() + pos * m_nFields + offset; }

    store_t &at(id_t id, std::size_t k = 0) { return rawAt(m_kernel->getRawIndex(id), k); }
    const store_t &at(id_t id, std::size_t k = 0) const { return rawAt(m_kernel->getRawIndex(id), k); }
    store_t &operator[](id_t id) { return rawAt(m_kernel->getRawIndex(id), 0); }
    const store_t &operator[](id_t id) const { return rawAt(m_kernel->getRawIndex(id), 0); }
    store_t *data(id_t id, std::size_t offset = 0) { return rawData(m_kernel->getRawIndex(id), offset); }
    const store_t *data(id_t id, std::size_t offset = 0) const { return rawData(m_kernel->getRawIndex(id), offset); }
    void set(id_t id, const value_t &value) { rawSet(m_kernel->getRawIndex(id), value); }
    void fill(const value_t &value) { m_data.assign(m_data.size(), (store_t) value); }

private:
    std::size_t m_nFields = 1;
    const PiercedKernel<id_t> *m_kernel = nullptr;
    std::vector<store_t> m_data;
};

} // namespace bitpit

#endif
