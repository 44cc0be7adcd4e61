// solver2d-b200 — index/revision pools (behaviour of reference src/pool.c:36-170).
// Slot reuse order matters to clients (ids) and to parity (slot order is the natural constraint order), so the
// semantics are kept: LIFO free list, revision incremented on reuse, growth to 1.5x with the new slots appended to an
// empty free list in ascending order.
#include "s2_host.h"

#include <stdlib.h>
#include <string.h>

static s2Object* s2SlotAt(s2Pool* pool, int32_t index)
{
	return (s2Object*)(pool->memory + (size_t)index * pool->objectSize);
}

// chain slots [first, last] into a free list ending in `tail`
static void s2ChainFree(s2Pool* pool, int32_t first, int32_t last, int32_t tail)
{
	for (int32_t i = first; i <= last; ++i)
	{
		s2Object* o = s2SlotAt(pool, i);
		o->index = i;
		o->next = i < last ? i + 1 : tail;
		o->revision = 0;
	}
}

s2Pool s2CreatePool(int32_t objectSize, int32_t capacity)
{
	s2Pool pool;
	pool.objectSize = objectSize;
	pool.capacity = capacity > 1 ? capacity : 1;
	pool.count = 0;
	pool.memory = (char*)calloc((size_t)pool.capacity, (size_t)objectSize);
	pool.freeList = 0;
	s2ChainFree(&pool, 0, pool.capacity - 1, S2_NULL_INDEX);
	return pool;
}

void s2DestroyPool(s2Pool* pool)
{
	free(pool->memory);
	memset(pool, 0, sizeof(*pool));
	pool->freeList = S2_NULL_INDEX;
}

s2Object* s2AllocObject(s2Pool* pool)
{
	if (pool->freeList != S2_NULL_INDEX)
	{
		s2Object* o = s2SlotAt(pool, pool->freeList);
		o->index = pool->freeList;
		o->revision += 1;
		pool->freeList = o->next;
		o->next = o->index;
		pool->count += 1;
		return o;
	}

	int32_t oldCapacity = pool->capacity;
	int32_t newCapacity = oldCapacity + oldCapacity / 2;
	newCapacity = newCapacity > 2 ? newCapacity : 2;
	char* memory = (char*)calloc((size_t)newCapacity, (size_t)pool->objectSize);
	memcpy(memory, pool->memory, (size_t)oldCapacity * pool->objectSize);
	free(pool->memory);
	pool->memory = memory;
	pool->capacity = newCapacity;

	s2Object* o = s2SlotAt(pool, oldCapacity);
	o->index = oldCapacity;
	o->revision = 0;
	o->next = o->index;

	if (oldCapacity + 1 <= newCapacity - 1)
	{
		pool->freeList = oldCapacity + 1;
		s2ChainFree(pool, oldCapacity + 1, newCapacity - 1, S2_NULL_INDEX);
	}
	else
	{
		pool->freeList = S2_NULL_INDEX;
	}
	pool->count += 1;
	return o;
}

void s2FreeObject(s2Pool* pool, s2Object* object)
{
	object->next = pool->freeList;
	pool->freeList = object->index;
	pool->count -= 1;
}
