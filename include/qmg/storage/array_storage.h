// quantum-mg on B200 -- ArrayStorageMG: check-out / check-in pool of equal-length DEVICE arrays
// (/root/reference/storage/array_storage.h:23-155).  The K-cycle takes every temporary from these pools, so no
// allocation happens inside a solve once the pools have grown to their steady size.
#ifndef QMG_B200_ARRAY_STORAGE
#define QMG_B200_ARRAY_STORAGE

#include <iostream>
#include <vector>
#include "blas/generic_vector.h"

template <typename T>
class ArrayStorageMG
{
private:
  ArrayStorageMG(ArrayStorageMG const&);
  ArrayStorageMG& operator=(ArrayStorageMG const&);

  struct Slot { T* ptr; bool out; };
  const long array_length;
  std::vector<Slot> slots;
  int n_checked;

public:
  ArrayStorageMG(const long length, const int n_prealloc = 1) : array_length(length), n_checked(0)
  {
    if (n_prealloc < 1) std::cout << "[QMG-ERROR]: ArrayStorageMG cannot preallocate less than one vector.\n";
    for (int i = 0; i < n_prealloc; i++) { Slot s = { allocate_vector<T>(array_length), false }; slots.push_back(s); }
  }
  ~ArrayStorageMG() { for (size_t i = 0; i < slots.size(); i++) deallocate_vector(&slots[i].ptr); }

  T* check_out()
  {
    n_checked++;
    for (size_t i = 0; i < slots.size(); i++) if (!slots[i].out) { slots[i].out = true; return slots[i].ptr; }
    Slot s = { allocate_vector<T>(array_length), true };
    slots.push_back(s);
    return s.ptr;
  }
  void check_in(T* arr)
  {
    for (size_t i = 0; i < slots.size(); i++)
    {
      if (slots[i].ptr != arr) continue;
      if (slots[i].out) { slots[i].out = false; n_checked--; }
      else std::cout << "[QMG_WARNING]: Returned array that wasn't checked out.\n";
      return;
    }
    std::cout << "[QMG_WARNING]: Returned array that doesn't live in library.\n";
  }
  int get_number_allocated() { return (int)slots.size(); }
  int get_number_checked() { return n_checked; }
  // release idle arrays beyond `minimum` (never the first one, like the reference)
  void consolidate(int minimum = 1)
  {
    for (int i = (int)slots.size() - 1; i > 0 && (int)slots.size() > minimum; i--)
      if (!slots[i].out) { deallocate_vector(&slots[i].ptr); slots.erase(slots.begin() + i); }
  }
};

#endif
