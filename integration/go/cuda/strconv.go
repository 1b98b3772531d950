package cuda

import "strconv"

func mustParseFloat(s string) float64 {
	x, err := strconv.ParseFloat(s, 64) // go/text/parse.go:163
	if err != nil {
		panic(err.Error())
	}
	return x
}
