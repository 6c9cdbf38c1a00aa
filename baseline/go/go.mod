module rpkbaseline

go 1.24
