/* placeholder, filled below */
