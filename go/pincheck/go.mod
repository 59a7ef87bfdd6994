module bloomsearch-b200/pincheck

go 1.21

// the versions the reference pins (reference go.mod:6,13)
require (
	github.com/bits-and-blooms/bitset v1.10.0
	github.com/bits-and-blooms/bloom/v3 v3.7.0
)
