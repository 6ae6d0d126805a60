module goref

go 1.20

require (
	github.com/consensys/gnark-crypto v0.11.0
	github.com/jsign/curdleproofs v0.0.0
)

// the unmodified reference checkout; override with `go mod edit -replace` for another location
replace github.com/jsign/curdleproofs => /root/reference
