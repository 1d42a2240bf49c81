module abdump

go 1.23.3

// the versions the reference pins (/root/reference/go.mod:6-7)
require (
	github.com/consensys/gnark v0.11.0
	github.com/consensys/gnark-crypto v0.14.1-0.20241217131346-b998989abdbe
)
