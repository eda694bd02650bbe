// frontier.hxx -- frontier_t<T>: the input/output of every traversal operator.
// Public surface of gunrock/src/frontier.hxx:12-99 (size/capacity/type/data, load from
// device or host, resize with the reference's overflow message).
#pragma once
#include "graph.hxx"

namespace gunrock {

enum frontier_type_t { edge_frontier = 0, node_frontier = 1 };

template <typename type_t>
class frontier_t {
    size_t _size = 0;
    size_t _capacity = 1;
    frontier_type_t _type = node_frontier;
    bool _hole_free = false;   // see hole_free()
    std::shared_ptr<mem_t<type_t>> _data;

    void overflow(const char *what, size_t wanted) const {
        std::printf("Overflow during frontier %s. Capacity is %d, size of the data to %s is %d.\n", what,
                    (int)_capacity, what[0] == 'l' ? "load" : "resize", (int)wanted);
        std::exit(0);   // reference behaviour (frontier.hxx:53-59,84-89); the C ABI returns B200_ERR_OVERFLOW instead
    }

   public:
    frontier_t() : _data(std::make_shared<mem_t<type_t>>()) {}
    frontier_t(context_t &context, size_t capacity, size_t size = 0, frontier_type_t type = node_frontier)
        : _size(size), _capacity(capacity), _type(type), _data(std::make_shared<mem_t<type_t>>(capacity, context)) {}
    frontier_t(const frontier_t &) = delete;
    frontier_t &operator=(const frontier_t &) = delete;
    frontier_t(frontier_t &&rhs) : frontier_t() { swap(rhs); }
    frontier_t &operator=(frontier_t &&rhs) {
        swap(rhs);
        return *this;
    }

    void swap(frontier_t &rhs) {
        std::swap(_size, rhs._size);
        std::swap(_capacity, rhs._capacity);
        std::swap(_type, rhs._type);
        std::swap(_hole_free, rhs._hole_free);
        _data.swap(rhs._data);
    }

    // device -> frontier
    cudaError_t load(mem_t<type_t> &target) {
        if (target.size() > _capacity) overflow("loading", target.size());
        _size = target.size();
        _hole_free = false;
        return dtod(_data->data(), target.data(), target.size());
    }
    // host -> frontier
    cudaError_t load(std::vector<type_t> target) {
        if (target.size() > _capacity) overflow("loading", target.size());
        _size = target.size();
        _hole_free = false;
        return htod(_data->data(), target);
    }
    void resize(size_t size) {
        if (size > _capacity) overflow("resizing", size);
        _size = size;
    }
    size_t capacity() const { return _capacity; }
    size_t size() const { return _size; }
    frontier_type_t type() const { return _type; }
    // Not in the reference: set by the operators that write a COMPACTED list (the default advance_forward_kernel, every
    // filter) and cleared by load().  A filter whose functor only drops the -1 holes of an un-compacted advance
    // (Functor::cond_filter_drops_only_holes) then has nothing to do but hand the list on.  A stale "true" after
    // somebody wrote holes through data() is harmless to the operators: they skip negative items.
    bool hole_free() const { return _hole_free; }
    void set_hole_free(bool v) { _hole_free = v; }
    std::shared_ptr<mem_t<type_t>> data() const { return _data; }
};

}  // namespace gunrock
